"""Synchronisation protocol of the persistent forward attention kernel (csrc/attn.cu: attn_pfwd2_kernel), checked on CPU
by a randomised discrete-event simulation.

The four roles (TMA producer, MMA-issuing warp, two softmax groups) are transcribed operation by operation — every
mbarrier wait with the parity expression the CUDA code uses, every commit / arrive, every buffer read and write — and run
under a random scheduler with asynchronous TMA completions and an in-order tensor pipe.  The simulation fails on
  * a deadlock (no agent can make progress before all have finished);
  * a parity wait that is not for the barrier's current or immediately preceding phase (on hardware: a wait that never
    returns, or one that returns one phase early);
  * a data hazard: an S / P ring slot, K/V stage, Q buffer or O accumulator read while it holds another block's or tile's
    data, or overwritten before its last reader is done.
What is specific to attn_pfwd2_kernel and modelled here: the MMA warp takes the waits of the NEXT S issue (Q buffer, K/V
stage) before it waits for P; the lazy-rescale path waits on the K/V stage's bar_empty of the group's previous block
(there is no separate "PV done" barrier); the two-group merge of tile t is DEFERRED behind each group's first block of
tile t+1, with the first PV of tile t+1 held back by bar_ofree.
Parameters sweep the cases that matter: 1, 2, 3 and many key blocks per tile (ring shorter / longer than a tile, a group
without blocks), 1 to 7 tiles per CTA (Q double buffer wrap-around), several random schedules each.
"""
import random

import pytest

STAGES = 5  # F3_STAGES


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than its count"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def ready(self, parity, want_phase):
        """try_wait.parity(parity): true once the phase with that parity has completed.  `want_phase` is the phase index
        the caller MEANS (bookkeeping of the test only): the hardware can only tell adjacent phases apart."""
        assert parity == (want_phase & 1), f"{self.name}: parity expression disagrees with the intended phase {want_phase}"
        assert self.phase in (want_phase, want_phase + 1), \
            f"{self.name}: waiting for phase {want_phase} while the barrier is in phase {self.phase} (not adjacent)"
        return self.phase == want_phase + 1


class Sim:
    def __init__(self, ntiles, nkb, seed):
        self.rnd = random.Random(seed)
        self.ntiles, self.nkb = ntiles, nkb
        B = Barrier
        self.bar_q = [B(f"bar_q{i}", 1) for i in range(2)]
        self.bar_qfree = [B(f"bar_qfree{i}", 1) for i in range(2)]
        self.bar_full = [B(f"bar_full{i}", 1) for i in range(STAGES)]
        self.bar_empty = [B(f"bar_empty{i}", 1) for i in range(STAGES)]
        self.bar_s = [B(f"bar_s{i}", 1) for i in range(3)]
        self.bar_p = [B(f"bar_p{i}", 1) for i in range(3)]      # 128 threads of ONE group -> modelled as one arrival
        self.bar_o = B("bar_o", 1)
        self.bar_ofree = B("bar_ofree", 2)                      # 256 threads = both groups -> two arrivals
        # buffer contents (what the data currently IS), for the hazard checks
        self.qbuf = [None, None]            # tile whose Q is in the buffer
        self.stage = [None] * STAGES        # global block whose K/V is in the stage
        self.ring = [None] * 3              # ("S", G) or ("P", G)
        self.o_acc = [(-1, 0), (-1, 0)]     # per group: (tile, number of PVs accumulated)
        self.async_q = []                   # pending asynchronous completions (TMA): callables
        self.mma_q = []                     # tensor pipe: callables executed strictly in issue order
        self.done_outputs = []
        self.sync_arrivals, self.o_read = {}, {}

    # ---- asynchronous engines
    def tma(self, fn):
        self.async_q.append(fn)

    def mma(self, fn):
        self.mma_q.append(fn)

    # ---- roles (generators yield a predicate to wait on, or None to just give up the time slice)
    def producer(self):
        s, ph = 0, 0
        G = 0
        for tl in range(self.ntiles):
            qb = tl & 1
            if tl >= 2:
                yield lambda qb=qb, tl=tl: self.bar_qfree[qb].ready(((tl >> 1) - 1) & 1, (tl >> 1) - 1)
                assert self.qbuf[qb] == tl - 2

            def land_q(qb=qb, tl=tl):
                self.qbuf[qb] = tl
                self.bar_q[qb].arrive()
            self.tma(land_q)
            for j in range(self.nkb):
                # bar_empty phase k is completed by the PV of the k-th block that used this stage; first pass is free
                use = G // STAGES
                if use > 0:
                    yield lambda s=s, ph=ph, use=use: self.bar_empty[s].ready(ph ^ 1, use - 1)
                else:
                    assert ph ^ 1 == 1  # fresh barrier: try_wait(parity 1) passes immediately

                def land_kv(s=s, G=G):
                    self.stage[s] = G
                    self.bar_full[s].arrive()
                self.tma(land_kv)
                G += 1
                s += 1
                if s == STAGES:
                    s, ph = 0, ph ^ 1
                yield None

    def mma_warp(self):
        st = dict(s_tl=0, s_j=0, s_stage=0, s_sph=0, s_buf=0, s_G=0)

        def wait_next_S():
            if st["s_tl"] >= self.ntiles:
                return
            tl, qb = st["s_tl"], st["s_tl"] & 1
            if st["s_j"] == 0:
                yield lambda: self.bar_q[qb].ready((tl >> 1) & 1, tl >> 1)
            stage, sph, G = st["s_stage"], st["s_sph"], st["s_G"]
            yield lambda: self.bar_full[stage].ready(sph, G // STAGES)

        def issue_next_S():
            if st["s_tl"] >= self.ntiles:
                return
            tl, qb = st["s_tl"], st["s_tl"] & 1
            stage, G, buf = st["s_stage"], st["s_G"], st["s_buf"]

            def exec_S(tl=tl, qb=qb, stage=stage, G=G, buf=buf):
                assert self.qbuf[qb] == tl, f"S({G}) reads Q buffer {qb} holding tile {self.qbuf[qb]}, wants {tl}"
                assert self.stage[stage] == G, f"S({G}) reads K/V stage {stage} holding block {self.stage[stage]}"
                assert self.ring[buf] is None or self.ring[buf] == ("consumed", G - 3), \
                    f"S({G}) overwrites ring slot {buf} = {self.ring[buf]}"
                self.ring[buf] = ("S", G)
            self.mma(exec_S)
            self.mma(lambda buf=buf: self.bar_s[buf].arrive())
            st["s_stage"] += 1
            if st["s_stage"] == STAGES:
                st["s_stage"], st["s_sph"] = 0, st["s_sph"] ^ 1
            st["s_buf"] = (st["s_buf"] + 1) % 3
            st["s_G"] += 1
            st["s_j"] += 1
            if st["s_j"] == self.nkb:
                self.mma(lambda qb=qb: self.bar_qfree[qb].arrive())
                st["s_j"] = 0
                st["s_tl"] += 1

        for _ in range(3):
            yield from wait_next_S()
            issue_next_S()
        buf, cs, ppar, G = 0, 0, 0, 0
        for tl in range(self.ntiles):
            for j in range(self.nkb):
                g = j & 1
                yield from wait_next_S()  # the next S issue's waits are taken BEFORE the wait for P
                yield lambda buf=buf, ppar=ppar, G=G: self.bar_p[buf].ready(ppar, G // 3)
                if j == 0 and tl > 0:
                    yield lambda tl=tl: self.bar_ofree.ready((tl - 1) & 1, tl - 1)

                def exec_PV(tl=tl, j=j, g=g, buf=buf, cs=cs, G=G):
                    assert self.ring[buf] == ("P", G), f"PV({G}) reads ring slot {buf} = {self.ring[buf]}"
                    assert self.stage[cs] == G, f"PV({G}) reads V stage {cs} holding block {self.stage[cs]}"
                    if j >= 2:
                        assert self.o_acc[g][0] == tl, f"PV({G}) accumulates into O_{g} of tile {self.o_acc[g][0]}"
                        self.o_acc[g] = (tl, self.o_acc[g][1] + 1)
                    else:
                        t_old, n_old = self.o_acc[g]
                        assert t_old < tl and (t_old < 0 or self.o_read.get(t_old, 0) == 2), \
                            f"first PV of tile {tl} overwrites O_{g} of tile {t_old} before both groups' merges read it"
                        self.o_acc[g] = (tl, 1)
                    self.ring[buf] = ("consumed", G)
                self.mma(exec_PV)
                self.mma(lambda cs=cs: self.bar_empty[cs].arrive())
                if j == self.nkb - 1:
                    self.mma(lambda: self.bar_o.arrive())
                issue_next_S()
                cs = (cs + 1) % STAGES
                buf += 1
                if buf == 3:
                    buf, ppar = 0, ppar ^ 1
                G += 1

    def group(self, g):
        def merge(tl):
            # wait for the tile's last PV, meet the other group (p2_groups_sync), THEN read both O accumulators
            yield lambda tl=tl: self.bar_o.ready(tl & 1, tl)
            self.sync_arrivals[tl] = self.sync_arrivals.get(tl, 0) + 1
            yield lambda tl=tl: self.sync_arrivals.get(tl, 0) == 2
            yield None  # the two groups do not read at the same instant
            n_own = [len(range(gg, self.nkb, 2)) for gg in (0, 1)]
            for gg in (0, 1):
                if n_own[gg]:
                    assert self.o_acc[gg] == (tl, n_own[gg]), \
                        f"group {g}: merge of tile {tl} reads O_{gg} = {self.o_acc[gg]}, wants {(tl, n_own[gg])}"
            self.o_read[tl] = self.o_read.get(tl, 0) + 1
            if self.o_read[tl] == 2:
                self.done_outputs.append(("merged", tl))
            self.bar_ofree.arrive()

        gbase = 0
        for tl in range(self.ntiles):
            kown = 0
            merged_prev = tl == 0
            for j in range(g, self.nkb, 2):
                G = gbase + j
                buf, spar = G % 3, (G // 3) & 1
                yield lambda buf=buf, spar=spar, G=G: self.bar_s[buf].ready(spar, G // 3)
                assert self.ring[buf] == ("S", G), f"group {g} reads ring slot {buf} = {self.ring[buf]}, wants S({G})"
                if kown > 0 and self.rnd.random() < 0.5:
                    # the lazy-rescale path: the PV of this group's previous block (G - 2) must be complete = its K/V stage free
                    Gp = G - 2
                    yield lambda Gp=Gp: self.bar_empty[Gp % STAGES].ready((Gp // STAGES) & 1, Gp // STAGES)
                    assert self.o_acc[g] == (tl, kown), f"rescale of O_{g}: {self.o_acc[g]} vs tile {tl}, {kown} PVs"
                yield None
                self.ring[buf] = ("P", G)
                self.bar_p[buf].arrive()
                kown += 1
                if not merged_prev:  # deferred merge of the previous tile, behind this tile's first own block
                    yield from merge(tl - 1)
                    merged_prev = True
            if not merged_prev:      # no own block in this tile (one key block per tile, group 1)
                yield from merge(tl - 1)
            gbase += self.nkb
        if self.ntiles > 0:
            yield from merge(self.ntiles - 1)

    # ---- scheduler
    def run(self):
        agents = {"producer": self.producer(), "mma": self.mma_warp(), "g0": self.group(0), "g1": self.group(1)}
        waiting = {k: None for k in agents}
        steps = 0
        while agents:
            steps += 1
            assert steps < 2_000_000, "runaway simulation"
            progressed = False
            choices = list(agents) + ["tma", "pipe"]
            self.rnd.shuffle(choices)
            for who in choices:
                if who == "tma":
                    if self.async_q:
                        self.async_q.pop(self.rnd.randrange(len(self.async_q)))()
                        progressed = True
                        break
                    continue
                if who == "pipe":
                    if self.mma_q:
                        self.mma_q.pop(0)()  # in order
                        progressed = True
                        break
                    continue
                if who not in agents:
                    continue
                pred = waiting[who]
                if pred is not None and not pred():
                    continue
                try:
                    waiting[who] = next(agents[who])
                except StopIteration:
                    del agents[who]
                progressed = True
                break
            if not progressed:
                # nothing runnable in this shuffle order: check whether ANYTHING could run
                if self.async_q or self.mma_q:
                    continue
                blocked = [k for k in agents if waiting[k] is not None and not waiting[k]()]
                assert len(blocked) < len(agents), f"deadlock: {blocked} all blocked, engines idle"
        while self.mma_q:
            self.mma_q.pop(0)()
        assert len([x for x in self.done_outputs if x[0] == "merged"]) == self.ntiles


@pytest.mark.parametrize("nkb", [1, 2, 3, 4, 8, 9])
@pytest.mark.parametrize("ntiles", [1, 2, 3, 5, 7])
def test_persistent_forward_protocol(ntiles, nkb):
    for seed in range(12):
        Sim(ntiles, nkb, seed).run()
