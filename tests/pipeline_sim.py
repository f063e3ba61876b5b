"""Discrete-event model of ONE CTA of qqq_gemm_kernel (test infrastructure, CPU only).

It replays the kernel's control flow — the warp roles of qqq_b200/csrc/qqq_gemm_sm100.cu with their rings, mbarriers,
arrival counts and PHASE-PARITY waits — on the schedule the host planner (qqq_b200_plan) produces, under randomised
latencies and warp interleavings, and checks what a GPU run can only show as a hang or a rare wrong result:

  * no deadlock: every role finishes;
  * every parity wait passes on exactly the phase it was meant for.  Hardware lets `try_wait.parity p` pass whenever the
    barrier's current phase bit differs from p, so a waiter that is a whole phase behind passes on stale data (the
    round-1 weight-ring bug: ring depth not a multiple of the unpack-group ownership period) and one that is a phase ahead
    hangs; the model asserts `completed_phases == intended_phase + 1` at every pass;
  * every 16-token chunk of every accumulator is drained exactly once.

Cross-CTA split-K waits are not modelled (publishers never wait, so they cannot close a cycle).  CTA pairs
(cta_group::2): `PairSim` runs both CTAs on one clock — own weight rings, unpack and epilogue warps, half of the token
tile each; the leader's MMA warp waits on barriers that collect both CTAs' arrivals and its commits are multicast to
both (`twin=True` is the cheaper approximation: one CTA with doubled arrival counts).
"""
from __future__ import annotations

import random

K_MAX_A_SLOTS = 8
K_WARPS = 20
K_UNPACK_WARP0 = 4


class MBar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.completed = name, count, count, 0

    def arrive(self, n=1):
        for _ in range(n):
            self.pending -= 1
            assert self.pending >= 0, f"{self.name}: more arrivals than the barrier's count in one phase"
            if self.pending == 0:
                self.completed += 1
                self.pending = self.count

    def passes(self, parity):
        return (self.completed & 1) != parity


class Ring:
    def __init__(self, n):
        self.n, self.idx, self.phase = n, 0, 0

    def advance(self):
        self.idx += 1
        if self.idx == self.n:
            self.idx, self.phase = 0, self.phase ^ 1


def segments(p, cta):
    """struct Sched (qqq_gemm_sm100.cu): [(tile, kb0, kb1)] of schedule index `cta`."""
    KU = p["k_units"]
    a_begin = min(cta * p["a_upc"], p["a_units"])
    a_end = min(a_begin + p["a_upc"], p["a_units"])
    segs = []
    if a_end > a_begin:
        for t in range(a_begin // KU, (a_end - 1) // KU + 1):
            segs.append((t, max(a_begin - t * KU, 0), min(KU, a_end - t * KU)))
    total = p["a_tiles"] + p["b_tiles"]
    segs += [(t, 0, KU) for t in range(p["a_tiles"] + cta, total, p["b_step"])]  # whole tiles: round-robin
    return segs


class Deadlock(AssertionError):
    pass


class CtaSim:
    def __init__(self, plan, cta, M, seed=0, dbuf_max_tok=208, twin=False, leader=None):
        """twin=True: CTA pair approximated by doubling this CTA's arrivals on the shared barriers.
        leader=<CtaSim>: this object is the PEER CTA (rank 1) of a faithfully modelled pair (see PairSim)."""
        self.p, self.rng = plan, (leader.rng if leader else random.Random(seed))
        self.leader, self.seed = leader, seed
        self.segs = segments(plan, cta)
        self.KU, self.KSUB, self.G = plan["k_units"], plan["ksub"], plan["unpack_groups"]
        self.NSW, self.NST = plan["stages_w"], plan["stages_t"]
        self.n_tok, self.M, self.m_tiles = plan["n_tok"], M, plan["m_tiles"]
        self.ndbuf = 2 if self.n_tok <= dbuf_max_tok else 1
        self.NA = min(K_MAX_A_SLOTS, (512 - self.ndbuf * self.n_tok) // (32 * self.KSUB))
        assert self.NA >= 1, "no room for the TMEM weight ring"
        self.n_epi = K_WARPS - K_UNPACK_WARP0 - 4 * self.G
        self.pair = 2 if twin else 1  # arrival multiplier on the leader's shared barriers
        mk = lambda name, n, cnt: [MBar(f"{name}[{i}]", cnt) for i in range(n)]  # noqa: E731
        self.fullw, self.emptyw = mk("fullw", self.NSW, 1), mk("emptyw", self.NSW, 4 * self.KSUB)
        self.fullt, self.emptyt = mk("fullt", self.NST, 1), mk("emptyt", self.NST, 1)
        self.afull, self.aempty = mk("afull", K_MAX_A_SLOTS, 4 * self.KSUB * self.pair), mk("aempty", K_MAX_A_SLOTS, 1)
        self.dfull, self.dempty = mk("dfull", 2, 1), mk("dempty", 2, self.n_epi * self.pair)
        self.clock = leader.clock if leader else {"t": 0, "seq": 0, "events": []}  # shared by the CTAs of a pair
        self.peer = None
        if leader is not None:
            # cta_group::2: the leader's MMA warp waits on ITS token-full / weights-unpacked / accumulator-empty barriers,
            # which collect both CTAs' arrivals; commits are multicast to the per-CTA barriers of both
            leader.peer = self
            leader.fullt = [MBar(f"fullt[{i}]", 2) for i in range(self.NST)]  # model: one arrival per CTA's TMA
            for b in leader.afull:
                b.count = b.pending = 4 * self.KSUB * 2
            for b in leader.dempty:
                b.count = b.pending = self.n_epi * 2
            self.fullt, self.afull, self.dempty = leader.fullt, leader.afull, leader.dempty
        self.drained = {}  # (segment, quadrant, chunk) -> count
        self.units = [(sg, kb) for sg, (_, kb0, kb1) in enumerate(self.segs) for kb in range(kb0, kb1)]

    # ---- asynchronous engines ---------------------------------------------------------------------------------
    @property
    def time(self):
        return self.clock["t"]

    def later(self, lo, hi, fn):
        self.clock["seq"] += 1
        self.clock["events"].append((self.time + self.rng.randint(lo, hi), self.clock["seq"], fn))

    def arrive_twice_if_pair(self, bar):
        bar.arrive(self.pair)  # the twin CTA does the same thing at (in the model) the same moment

    # ---- roles (generators yielding ("wait", bar, parity, intended_phase) | ("sleep", cycles)) -------------------
    def weights_producer(self):
        st = Ring(self.NSW)
        for n, _ in enumerate(self.units):
            if n >= self.NSW:  # the first ring of stages is requested without waiting
                yield ("wait", self.emptyw[st.idx], st.phase ^ 1, n // self.NSW - 1)
            bar = self.fullw[st.idx]
            self.later(200, 3000 if self.rng.random() < 0.05 else 900, bar.arrive)  # DRAM / L2, sometimes very late
            yield ("sleep", self.rng.randint(1, 40))
            st.advance()

    def tokens_producer(self):
        st = Ring(self.NST)
        for n, _ in enumerate(self.units):
            yield ("wait", self.emptyt[st.idx], st.phase ^ 1, n // self.NST - 1)
            self.later(150, 700, self.fullt[st.idx].arrive)
            yield ("sleep", self.rng.randint(1, 40))
            st.advance()

    def mma_issuer(self):
        st, as_ = Ring(self.NST), Ring(self.NA)
        n = 0
        last_retire = [0]
        for sg, (_, kb0, kb1) in enumerate(self.segs):
            dbuf, use = sg % self.ndbuf, sg // self.ndbuf
            yield ("wait", self.dempty[dbuf], (use & 1) ^ 1, use - 1)
            for kb in range(kb0, kb1):
                yield ("wait", self.fullt[st.idx], st.phase, n // self.NST)
                yield ("wait", self.afull[as_.idx], as_.phase, n // self.NA)
                # tcgen05.commit: the arrivals happen when the MMAs retire, in issue order
                due = max(last_retire[0], self.time) + self.rng.randint(100, 600)
                last_retire[0] = due
                bars = [self.aempty[as_.idx], self.emptyt[st.idx]] + ([self.dfull[dbuf]] if kb == kb1 - 1 else [])
                if self.peer is not None:  # tcgen05.commit.multicast::cluster: the same barriers in the peer CTA
                    bars += [self.peer.aempty[as_.idx], self.peer.emptyt[st.idx]] + (
                        [self.peer.dfull[dbuf]] if kb == kb1 - 1 else [])
                self.clock["seq"] += 1
                self.clock["events"].append((due, self.clock["seq"], lambda bars=bars: [b.arrive() for b in bars]))
                yield ("sleep", self.rng.randint(20, 200))
                st.advance()
                as_.advance()
                n += 1

    def seg_rows(self, sg):
        tile = self.segs[sg][0]
        mt = tile % self.m_tiles
        return min(self.n_tok, self.M - mt * self.n_tok)

    def whole(self, sg):
        _, kb0, kb1 = self.segs[sg]
        return kb0 == 0 and kb1 == self.KU

    def finisher(self, sg):
        """Whether this CTA finishes the tile of segment sg.  For split tiles that is decided by arrival order across
        CTAs (ticket); the model draws it once per segment so that every role of the CTA sees the same answer."""
        return self.whole(sg) or random.Random(hash((id(self.p) & 0xFFFF, self.seed, sg))).random() < 0.5

    def drain(self, sg, q, first_chunk, step):
        rows = self.seg_rows(sg)
        for mb in range(16 * first_chunk, rows, 16 * step):
            key = (sg, q, mb // 16)
            self.drained[key] = self.drained.get(key, 0) + 1
            yield ("sleep", self.rng.randint(300, 900))

    def unpack_warp(self, grp, q):
        st, as_ = Ring(self.NSW), Ring(self.NA)
        turn = 0
        for u in range(len(self.units)):
            stage_ready = slot_ready = False
            for _sub in range(self.KSUB):
                if turn == grp:
                    if not stage_ready:
                        yield ("wait", self.fullw[st.idx], st.phase, u // self.NSW)
                        stage_ready = True
                    if not slot_ready:
                        yield ("wait", self.aempty[as_.idx], as_.phase ^ 1, u // self.NA - 1)
                        slot_ready = True
                    yield ("sleep", self.rng.randint(60, 500))
                    self.afull[as_.idx].arrive(self.pair)
                    self.emptyw[st.idx].arrive()
                turn = 0 if turn == self.G - 1 else turn + 1
            st.advance()
            as_.advance()

    def epilogue_warp(self, e):
        q, eh = e % 4, e // 4
        for sg in range(len(self.segs)):
            dbuf, use = sg % self.ndbuf, sg // self.ndbuf
            yield ("wait", self.dfull[dbuf], use & 1, use)
            if not self.whole(sg):
                if e == 0:
                    # ticket; the finisher also waits until the other contributors' partials are published (other CTAs,
                    # never blocked by this one)
                    yield ("sleep", self.rng.randint(20, 2000 if self.finisher(sg) else 100))
                    self.ticket_ready = sg
                else:  # named barrier of the epilogue warps around the ticket
                    while getattr(self, "ticket_ready", -1) < sg:
                        yield ("sleep", self.rng.randint(5, 50))
            yield from self.drain(sg, q, eh, self.n_epi // 4)
            self.dempty[dbuf].arrive(self.pair)

    # ---- scheduler ------------------------------------------------------------------------------------------------
    def roles(self, tag=""):
        roles = {tag + "W": self.weights_producer(), tag + "T": self.tokens_producer()}
        if self.leader is None:
            roles[tag + "M"] = self.mma_issuer()  # pair: the leader CTA only
        for g in range(self.G):
            for q in range(4):
                roles[f"{tag}U{g}.{q}"] = self.unpack_warp(g, q)
        for e in range(self.n_epi):
            roles[f"{tag}E{e}"] = self.epilogue_warp(e)
        return roles

    def run(self, extra_roles=None):
        roles = self.roles()
        if extra_roles:
            roles.update(extra_roles)
        clock = self.clock
        state = {k: ("ready", None) for k in roles}  # ready | ("wait", bar, parity, phase) | ("sleep", until)
        done = set()
        while len(done) < len(roles):
            # fire asynchronous completions that are due
            clock["events"].sort()
            while clock["events"] and clock["events"][0][0] <= clock["t"]:
                clock["events"].pop(0)[2]()
            progressed = False
            names = [k for k in roles if k not in done]
            self.rng.shuffle(names)
            for k in names:
                kind, arg = state[k]
                if kind == "sleep" and arg > clock["t"]:
                    continue
                if kind == "wait":
                    bar, parity, phase = arg
                    if not bar.passes(parity):
                        continue
                    assert bar.completed == phase + 1, (
                        f"{k}: wait on {bar.name} for phase {phase} passed with {bar.completed} phases completed "
                        f"({'stale parity' if bar.completed < phase + 1 else 'barrier ran ahead'})")
                try:
                    ev = next(roles[k])
                except StopIteration:
                    done.add(k)
                    progressed = True
                    continue
                progressed = True
                if ev[0] == "sleep":
                    state[k] = ("sleep", clock["t"] + ev[1])
                else:
                    state[k] = ("wait", ev[1:])
            if not progressed:
                sleepers = [state[k][1] for k in names if state[k][0] == "sleep"]
                nxt = min(sleepers + [e[0] for e in clock["events"]], default=None)
                if nxt is None:
                    blocked = {k: (state[k][1][0].name, state[k][1][2]) for k in names if state[k][0] == "wait"}
                    raise Deadlock(f"deadlock at t={clock['t']}: {blocked}")
                clock["t"] = max(clock["t"] + 1, nxt)
        self.check_drained()
        if self.peer is not None:
            self.peer.check_drained()
        return clock["t"]

    def check_drained(self):
        # every chunk of every segment exactly once per quadrant
        for sg in range(len(self.segs)):
            for q in range(4):
                for c in range(-(-self.seg_rows(sg) // 16)):
                    assert self.drained.get((sg, q, c), 0) == 1, f"segment {sg} quadrant {q} chunk {c}: drained {self.drained.get((sg, q, c), 0)}x"


def PairSim(plan, cta, M, seed=0, dbuf_max_tok=208):
    """Both CTAs of a pair (cluster of 2, cta_group::2) on one clock: each has its own weight ring and unpack / epilogue
    warps and loads half of the token tile; the leader's MMA warp consumes both and its commits release both."""
    lead = CtaSim(plan, cta, M, seed=seed, dbuf_max_tok=dbuf_max_tok)
    peer = CtaSim(plan, cta, M, seed=seed, dbuf_max_tok=dbuf_max_tok, leader=lead)
    return lead.run(extra_roles=peer.roles("peer:"))
