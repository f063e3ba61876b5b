"""Print selected raw metrics from an .ncu-rep (dev tooling)."""
import csv,sys,subprocess
f=sys.argv[1]
out=subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=sys.argv[2:]
for i,h in enumerate(hdr):
    if h in want or any(w.endswith('*') and h.startswith(w[:-1]) for w in want): print(f'  {h:75s} {units[i]:14s} {vals[i]}')
