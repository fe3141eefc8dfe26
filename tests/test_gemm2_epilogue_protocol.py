"""Synchronisation protocol of the GEMM kernel's epilogue (csrc/gemm2.cu: gemm2_kernel, round-2 version with two epilogue
warp groups), checked on CPU by a randomised discrete-event simulation.

The MMA-issuing warp and every epilogue warp (4 per group) are transcribed operation by operation: the accumulator
hand-shake (bar_acc_full / bar_acc_empty with the EARLY relaxed release after a warp's last TMEM read), the per-group named
barriers and the cross-group one, the staging buffers with their asynchronous TMA stores (`wait_group.read N` semantics),
the residual-tile TMA loads (prefetched before the accumulator wait / one chunk ahead in single-group mode, at the chunk
start in two-group mode) with bar_res phases, the bias columns staged through a shared-memory double buffer, the GEGLU mode
(both groups on one chunk, three staging tiles) and the GEGLU-backward mode (u tiles in by TMA, two tiles per group, results
in place).  The simulation fails on
  * a deadlock;
  * an mbarrier parity wait that is not for the barrier's current or immediately preceding phase;
  * a data hazard: an accumulator buffer overwritten while a warp still has to read it, or read while it holds another
    tile; a staging buffer written while a TMA store is still reading it or before its residual tile arrived, or refilled by
    TMA while a thread still uses it; a store issued before every thread of its group wrote; bias columns read from a buffer
    that holds another tile's, or overwritten while somebody still reads them.
Mutations of the protocol (release before the last read, no drain wait, the single-group drain depth in two-group mode,
no cross-group barrier after the bias staging) must each be caught — otherwise the checks would be vacuous.
Reference behaviour this kernel replaces: the epilogues of cuBLASLt / cuDNN behind diffusers' Linear / Conv2d layers.
"""
import random

import pytest


class Barrier:
    """mbarrier with phase bookkeeping (see tests/test_attn_persistent_protocol.py)."""

    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than its count"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def ready(self, parity, want_phase):
        assert parity == (want_phase & 1), f"{self.name}: parity expression disagrees with the intended phase {want_phase}"
        assert self.phase in (want_phase, want_phase + 1), \
            f"{self.name}: waiting for phase {want_phase} while the barrier is in phase {self.phase} (not adjacent)"
        return self.phase == want_phase + 1


class NamedBarrier:
    """bar.sync id, count: generation counter; an agent arrives and waits for its generation to complete."""

    def __init__(self, count):
        self.count, self.n, self.gen = count, 0, 0

    def arrive(self):
        g = self.gen
        self.n += 1
        if self.n == self.count:
            self.n = 0
            self.gen += 1
        return g


class Sim:
    def __init__(self, mode, groups, cols, res, prefetch, bias, seed, mut=None):
        self.rnd = random.Random(seed)
        self.mode, self.groups, self.two = mode, groups, groups == 2
        self.cols, self.res, self.prefetch, self.bias, self.mut = cols, res, prefetch, bias, mut
        self.ntiles = len(cols)
        self.acc_full = [Barrier(f"acc_full{b}", 1) for b in range(2)]
        self.acc_empty = [Barrier(f"acc_empty{b}", 4 * groups) for b in range(2)]
        self.bar_res = [Barrier(f"bar_res{b}", 1) for b in range(2)]
        self.gbars = [NamedBarrier(4) for _ in range(2)]
        self.allb = NamedBarrier(4 * groups)
        # TMEM accumulators: which tile they hold, and which warps still have columns of it to read
        self.acc_tile = [None, None]
        self.acc_readers = [set(), set()]
        # staging buffers (up to 4): data = what is in it, unread = TMA stores still reading it, users = warps inside a chunk on it
        self.st_data = [None] * 4
        self.st_unread = [0] * 4
        self.st_users = [set() for _ in range(4)]
        self.st_written = [set() for _ in range(4)]
        self.groups_pending = {}            # issuing warp -> list of [done?] per committed bulk group, in issue order
        self.bias_tile = [None, None]
        self.bias_readers = [dict(), dict()]   # warp -> tile whose bias columns it is still going to read
        self.async_q = []
        self.stored = []

    # ---------------------------------------------------------------- asynchronous engines
    def later(self, fn):
        self.async_q.append(fn)

    def store(self, who, bufs, what):
        for b in bufs:
            assert self.st_written[b] == self.st_users[b] and len(self.st_users[b]) > 0, \
                f"store of {what} from staging {b} before every thread of its group wrote"
            self.st_unread[b] += 1
        grp = [False]
        self.groups_pending.setdefault(who, []).append(grp)

        def read_done():
            for b in bufs:
                self.st_unread[b] -= 1
            grp[0] = True
            self.stored.append(what)
        self.later(read_done)

    def drained(self, who, n):
        g = self.groups_pending.get(who, [])
        return all(x[0] for x in (g[:-n] if n else g))

    def load_residual(self, b, what):
        assert self.st_unread[b] == 0, f"residual TMA load into staging {b} while a store is still reading it"
        assert not self.st_users[b], f"residual TMA load into staging {b} while {self.st_users[b]} still use it"

        def arrived():
            self.st_data[b] = ("res", what)
            self.bar_res[b % 2].arrive()
        self.later(arrived)

    # ---------------------------------------------------------------- roles
    def mma(self):
        for t in range(self.ntiles):
            buf, use = t & 1, t >> 1
            yield lambda: self.acc_empty[buf].ready((use & 1) ^ 1, use - 1) if use > 0 else True
            assert not self.acc_readers[buf], f"MMA overwrites accumulator {buf} while {self.acc_readers[buf]} still read tile {self.acc_tile[buf]}"
            self.acc_tile[buf] = ("partial", t)
            for _ in range(self.rnd.randint(1, 4)):
                yield None
            self.acc_tile[buf] = t
            self.acc_readers[buf] = {(g, w) for g in range(self.groups) for w in range(4)}
            self.acc_full[buf].arrive()

    def bar(self, nb):
        g = nb.arrive()
        return lambda: nb.gen > g

    def read_acc(self, me, buf, t, last):
        assert self.acc_tile[buf] == t, f"{me} reads accumulator {buf} holding {self.acc_tile[buf]}, wants tile {t}"
        assert me in self.acc_readers[buf], f"{me} reads accumulator {buf} after handing it back"
        if last:
            self.acc_readers[buf].discard(me)
            self.acc_empty[buf].arrive()

    def sts(self, me, b, chunk, need_res):
        assert self.st_unread[b] == 0, f"{me} writes staging {b} while a TMA store is still reading it"
        if need_res:
            assert self.st_data[b] == ("res", chunk) or self.st_data[b] == ("out", chunk), \
                f"{me} adds the residual of {chunk} but staging {b} holds {self.st_data[b]}"
        self.st_written[b].add(me)

    def epi(self, grp, w):
        me = (grp, w)
        lead = w == 0
        two, mode = self.two, self.mode
        c_first, c_step = (grp * 64, 128) if two else (0, 64)
        chunk_i, res_uses = 0, [0, 0]
        for t in range(self.ntiles):
            ncols, buf, use, tb = self.cols[t], t & 1, t >> 1, t & 1
            # ---- bias columns -> shared memory (group 0), cross-group barrier in two-group mode
            if self.bias:
                if grp == 0:
                    stale = {r: v for r, v in self.bias_readers[tb].items() if v != t}
                    assert not stale, f"bias buffer {tb} overwritten for tile {t} while {stale} still read an older tile's columns"
                    self.bias_tile[tb] = t
                if two and self.mut != "no_bias_barrier":
                    yield self.bar(self.allb)
                self.bias_readers[tb][me] = t
            plain = mode in ("plain", "gbwd")
            first_res_early = (mode == "plain" and self.res and self.prefetch and c_first < ncols)
            if first_res_early and lead:
                sb = grp if two else chunk_i & 1
                if self.mut != "no_drain":
                    n = 0 if two else 1
                    yield lambda n=n: self.drained(me, n)
                self.load_residual(sb, (t, c_first))
            yield lambda: self.acc_full[buf].ready(use & 1, use)
            if plain and c_first >= ncols:
                self.read_acc(me, buf, t, True)   # no chunk of this tile: hand the buffer back at once
            if mode == "geglu":
                for c0 in (0, 64):
                    if lead and grp == 0 and self.mut != "no_drain":
                        yield lambda: self.drained(me, 0)
                    yield self.bar(self.allb)
                    for b in (0, 1, 2):
                        self.st_users[b].add(me)
                    halves = (grp,) if two else (0, 1)
                    for hf in halves:
                        self.read_acc(me, buf, t, c0 == 64 and (two or hf == 1))
                        yield None
                        if self.bias:
                            assert self.bias_tile[tb] == t, f"{me} reads bias of tile {self.bias_tile[tb]} for tile {t}"
                        for b in (0, 1, 2):
                            self.sts(me, b, (t, c0), False)
                    if c0 == 64 and self.bias:
                        self.bias_readers[tb].pop(me, None)
                    yield self.bar(self.allb)
                    if lead and grp == 0:
                        self.store(me, (0, 1, 2), ("geglu", t, c0))
                        for b in (0, 1, 2):
                            self.st_users[b].clear()
                            self.st_written[b].clear()
                continue
            if mode == "gbwd":
                assert two
                hb, gb = grp * 2, grp * 2 + 1
                for c0 in range(c_first, ncols, c_step):
                    if lead:
                        if self.mut != "no_drain":
                            yield lambda: self.drained(me, 0)
                        for b in (hb, gb):
                            assert self.st_unread[b] == 0 and not self.st_users[b], f"u-tile load into busy staging {b}"
                        def arrived(t=t, c0=c0):
                            self.st_data[hb] = ("res", (t, c0))
                            self.st_data[gb] = ("res", (t, c0))
                            self.bar_res[grp].arrive()
                        self.later(arrived)
                    yield lambda: self.bar_res[grp].ready(res_uses[grp] & 1, res_uses[grp])
                    res_uses[grp] += 1
                    self.st_users[hb].add(me)
                    self.st_users[gb].add(me)
                    self.read_acc(me, buf, t, c0 + c_step >= ncols)
                    yield None
                    self.sts(me, hb, (t, c0), True)
                    self.sts(me, gb, (t, c0), True)
                    yield self.bar(self.gbars[grp])
                    if lead:
                        self.st_data[hb] = self.st_data[gb] = ("out", (t, c0))
                        self.store(me, (hb, gb), ("gbwd", t, c0))
                        for b in (hb, gb):
                            self.st_users[b].clear()
                            self.st_written[b].clear()
                continue
            # ---- plain mode: this group's chunks (full 64-column chunks; the direct-store path of a narrow chunk has no protocol)
            for c0 in range(c_first, ncols, c_step):
                sbuf = grp if two else chunk_i & 1
                if lead:
                    if self.res and (two or not self.prefetch):
                        if not (first_res_early and c0 == c_first):
                            if self.mut != "no_drain":
                                n = 0 if two else 1
                                if self.mut == "drain_depth_1_two_groups" and two:
                                    n = 1
                                yield lambda n=n: self.drained(me, n)
                            self.load_residual(sbuf, (t, c0))
                    elif self.res:
                        if c0 + 64 < ncols:
                            yield lambda: self.drained(me, 0)
                            self.load_residual(sbuf ^ 1, (t, c0 + 64))
                    elif self.mut != "no_drain":
                        n = 0 if two else 1
                        if self.mut == "drain_depth_1_two_groups" and two:
                            n = 1
                        yield lambda n=n: self.drained(me, n)
                yield self.bar(self.gbars[grp])
                if self.res:
                    u_ = res_uses[sbuf]
                    yield lambda u_=u_, sbuf=sbuf: self.bar_res[sbuf].ready(u_ & 1, u_)
                    res_uses[sbuf] += 1
                self.st_users[sbuf].add(me)
                last = c0 + c_step >= ncols
                if self.mut == "early_release" and c0 == c_first and not last:
                    self.read_acc(me, buf, t, True)       # hands the buffer back although a later chunk still has to read it
                else:
                    self.read_acc(me, buf, t, last)
                yield None
                if self.bias:
                    assert self.bias_tile[tb] == t, f"{me} reads bias of tile {self.bias_tile[tb]} for tile {t}"
                    if last:
                        self.bias_readers[tb].pop(me, None)
                self.sts(me, sbuf, (t, c0), self.res)
                yield self.bar(self.gbars[grp])
                if lead:
                    self.st_data[sbuf] = ("out", (t, c0))
                    self.store(me, (sbuf,), ("plain", t, c0))
                    self.st_users[sbuf].clear()
                    self.st_written[sbuf].clear()
                chunk_i += 1
            if self.bias and c_first >= ncols:
                self.bias_readers[tb].pop(me, None)

    # ---------------------------------------------------------------- scheduler
    def run(self):
        agents = [self.mma()] + [self.epi(g, w) for g in range(self.groups) for w in range(4)]
        waiting = [None] * len(agents)
        alive = [True] * len(agents)
        idle = 0
        while any(alive):
            progressed = False
            order = list(range(len(agents)))
            self.rnd.shuffle(order)
            for i in order:
                if not alive[i]:
                    continue
                if waiting[i] is not None and not waiting[i]():
                    continue
                try:
                    waiting[i] = next(agents[i])
                except StopIteration:
                    alive[i] = False
                progressed = True
                if self.rnd.random() < 0.5:
                    break
            if self.async_q and (not progressed or self.rnd.random() < 0.4):
                self.async_q.pop(self.rnd.randrange(len(self.async_q)))()
                progressed = True
            idle = 0 if progressed else idle + 1
            assert idle < 50, "deadlock: no agent can make progress"
        while self.async_q:
            self.async_q.pop()()


def expected_stores(mode, cols):
    out = []
    for t, nc in enumerate(cols):
        if mode == "geglu":
            out += [("geglu", t, 0), ("geglu", t, 64)]
        else:
            out += [(mode, t, c0) for c0 in range(0, nc, 64)]
    return sorted(out)


COLS = [[256], [320], [64], [128, 64, 256], [256, 256, 256, 256, 256], [64, 64, 64, 64], [320, 320], [192, 256, 128, 64, 320]]


@pytest.mark.parametrize("groups", [1, 2])
@pytest.mark.parametrize("res,prefetch", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("bias", [False, True])
def test_plain_epilogue_protocol(groups, res, prefetch, bias):
    for cols in COLS:
        for seed in range(12):
            s = Sim("plain", groups, cols, res, prefetch, bias, seed)
            s.run()
            assert sorted(s.stored) == expected_stores("plain", cols), (cols, seed)


@pytest.mark.parametrize("groups", [1, 2])
@pytest.mark.parametrize("bias", [False, True])
def test_geglu_epilogue_protocol(groups, bias):
    for ntiles in (1, 2, 3, 6):
        for seed in range(12):
            s = Sim("geglu", groups, [256] * ntiles, False, True, bias, seed)
            s.run()
            assert sorted(s.stored) == expected_stores("geglu", [256] * ntiles)


def test_geglu_backward_epilogue_protocol():
    for cols in ([256], [128], [256, 256, 128], [256] * 5):
        for seed in range(12):
            s = Sim("gbwd", 2, cols, True, True, False, seed)
            s.run()
            assert sorted(s.stored) == expected_stores("gbwd", cols)


@pytest.mark.parametrize("mut,mode,groups,res,bias", [
    ("early_release", "plain", 2, False, False),          # accumulator handed back before the warp's last TMEM read
    ("early_release", "plain", 1, False, False),
    ("no_drain", "plain", 2, False, False),               # staging buffer rewritten while the previous TMA store reads it
    ("no_drain", "plain", 1, False, False),
    ("no_drain", "plain", 2, True, False),                # residual load into a buffer a store still reads
    ("no_drain", "geglu", 2, False, False),
    ("no_drain", "gbwd", 2, True, False),
    ("drain_depth_1_two_groups", "plain", 2, False, False),  # wait_group.read 1 is right for two alternating buffers, not for one
    ("no_bias_barrier", "plain", 2, False, True),         # group 1 reads bias columns group 0 has not staged yet
])
def test_mutations_are_caught(mut, mode, groups, res, bias):
    caught = 0
    for cols in ([256, 256, 256, 256], [320, 320, 320], [128, 256, 256, 128, 256]):
        for seed in range(40):
            try:
                Sim(mode, groups, cols if mode != "geglu" else [256] * 4, res, True, bias, seed, mut=mut).run()
            except AssertionError:
                caught += 1
    assert caught > 0, f"mutation {mut} ({mode}, {groups} group(s)) was never detected"
