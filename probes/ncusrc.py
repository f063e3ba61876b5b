"""Top stall sites from an ncu source page CSV (dev tooling). usage: ncusrc.py rep [topN]"""
import csv,sys,subprocess,io
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
lines=out.splitlines()
start=[i for i,l in enumerate(lines) if l.startswith('"Address"')][0]
rd=csv.DictReader(io.StringIO("\n".join(lines[start:])))
rows=list(rd)
def f(r,k):
    try: return float(r[k])
    except: return 0.0
tot=sum(f(r,'# Samples') for r in rows)
print("total samples",tot)
rows2=sorted(rows,key=lambda r:-f(r,'# Samples'))[:topn]
stalls=[k for k in rows[0].keys() if k.startswith('stall_') and 'Not Issued' not in k]
for r in rows2:
    s={k:f(r,k) for k in stalls}
    top=sorted(s.items(),key=lambda kv:-kv[1])[:2]
    print(f"{f(r,'# Samples'):8.0f} {100*f(r,'# Samples')/tot:5.1f}%  {r['Address'][-5:]} {r['Source'][:90]:90s} {top[0][0]}={top[0][1]:.0f} {top[1][0]}={top[1][1]:.0f} exec={r['Instructions Executed']}")
