"""Host logic of the data-parallel gradient exchange (sdxl_training_improvements_b200/dp.py) on CPU:
  * shard_plan: reduce-scatter ownership covers every piece exactly once, 16-byte granular, staging offsets disjoint;
    a numpy simulation of reduce-scatter + all-gather over that plan leaves the exact sum on every rank;
  * plan_chunks: every parameter lands in exactly one chunk whose cut is at or after the last backward-tape position that
    writes its gradient; chunks tile the flat buffer; small (1-D) parameters are flushed with their chunk;
  * world-size-2 gloo: `try_create_exchange` declines collectively without CUDA (-> the NCCL / gloo all-reduce path).
The transport itself (IPC-mapped peer buffers, copy engines, flag words) needs >= 2 GPUs: tests/test_gpu_dp_exchange.py.
"""
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.unet_sdxl import tiny_config  # noqa: E402
from sdxl_training_improvements_b200 import dp  # noqa: E402
from sdxl_training_improvements_b200.params import SDXL_BASE, ParamStore  # noqa: E402


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_shard_plan_covers_every_piece_once(world):
    ranges = [(0, 8), (64, 1024), (4096, 8 * 37), (100000, 8 * 1001), (200000, 16)]
    owned = {}
    for r in range(world):
        offs, lens, soffs, used = dp.shard_plan(ranges, world, r)
        assert len(offs) == len(ranges)
        prev_end = 0
        for (off, n), so, sl, st in zip(ranges, offs, lens, soffs):
            assert so % 8 == 0 and sl % 8 == 0 and st % 8 == 0 and sl >= 0
            assert off <= so and so + sl <= off + n
            assert st >= prev_end  # staging regions of consecutive pieces do not overlap
            prev_end = st + sl
            for e in range(so, so + sl, 8):
                assert e not in owned, "two ranks own the same 16 bytes"
                owned[e] = r
        assert used >= prev_end
    want = {e for off, n in ranges for e in range(off, off + n, 8)}
    assert set(owned) == want


@pytest.mark.parametrize("world", [2, 4, 8])
def test_simulated_exchange_leaves_the_sum_everywhere(world):
    rng = np.random.default_rng(0)
    total = 8 * 5000
    ranges = [(0, 8 * 1200), (8 * 1200, 8 * 7), (8 * 3000, 8 * 2000)]  # a chunk need not cover the whole buffer
    bufs = [rng.integers(-8, 8, total).astype(np.float32) for _ in range(world)]
    exact = sum(bufs)
    plans = [dp.shard_plan(ranges, world, r) for r in range(world)]
    slot = dp.staging_slot_elems(total, world)
    # reduce-scatter: rank r pulls its shard from every peer into staging slot j and sums
    reduced = [b.copy() for b in bufs]
    for r in range(world):
        offs, lens, soffs, used = plans[r]
        assert used <= slot
        staging = np.zeros((world - 1, slot), np.float32)
        for j in range(world - 1):
            p = (r + 1 + j) % world
            for o, n, s in zip(offs, lens, soffs):
                staging[j, s:s + n] = bufs[p][o:o + n]
        for o, n, s in zip(offs, lens, soffs):
            reduced[r][o:o + n] = bufs[r][o:o + n] + staging[:, s:s + n].sum(0)
    # all-gather: rank r pushes its reduced shard into every peer
    final = [b.copy() for b in reduced]
    for r in range(world):
        offs, lens, _, _ = plans[r]
        for p in range(world):
            for o, n in zip(offs, lens):
                final[p][o:o + n] = reduced[r][o:o + n]
    inside = np.zeros(total, bool)
    for o, n in ranges:
        inside[o:o + n] = True
    for r in range(world):
        assert np.array_equal(final[r][inside], exact[inside])
        assert np.array_equal(final[r][~inside], bufs[r][~inside])  # outside the chunk nothing moves


def _synthetic_log(store, n_tape, seed=0):
    """A backward pass that writes parameter gradients roughly in reverse buffer order (like the UNet's), touches fused
    views that span several parameters, and writes a few parameters more than once."""
    rnd = random.Random(seed)
    lay = dp._layout(store)
    log = []
    big = [(o, n, name) for o, n, name in lay if name not in store.small_off]
    for i, (o, n, name) in enumerate(reversed(big)):
        pos = min(n_tape - 1, int(i * (n_tape - 1) / max(len(big) - 1, 1)) + rnd.randint(0, 2))
        log.append((pos, "flat", o, store._numel[name]))
        if rnd.random() < 0.1:  # a second, later write (e.g. the grouped K/V weight gradient at the very end)
            log.append((n_tape - 1, "flat", o, store._numel[name]))
    for name, so in store.small_off.items():
        pos = rnd.randint(0, n_tape - 1)
        log.append((pos, "small", so, store._numel[name]))
    return log


@pytest.mark.parametrize("cfg,target", [(tiny_config(), 4), (tiny_config(), 10)])
def test_plan_chunks_invariants(cfg, target):
    store = ParamStore(cfg, device="cpu")
    n_tape = 400
    log = _synthetic_log(store, n_tape)
    plan = dp.plan_chunks(store, log, n_tape, target_chunks=target)
    last = dp.last_touch_positions(store, log)
    assert plan.cuts == sorted(set(plan.cuts)) and plan.cuts[-1] == n_tape - 1
    assert 1 <= plan.n_chunks <= target + 2  # greedy cuts + the tail + one late cut that keeps the tail small
    # every parameter in exactly one chunk, final before its cut
    for name in store.offsets:
        k = plan.param_chunk[name]
        assert plan.cuts[k] >= last[name]
        if k > 0:
            assert plan.cuts[k - 1] < last[name], "parameter could have gone to an earlier chunk"
    # chunks tile the flat buffer
    cover = np.zeros(store.total, np.int32)
    for k, rg in enumerate(plan.ranges):
        for off, n in rg:
            assert off % 8 == 0 and n % 8 == 0
            cover[off:off + n] += 1
    assert (cover == 1).all()
    # every small parameter is flushed exactly once, with the chunk that owns its slot in the flat buffer
    seen = set()
    for k, segs in enumerate(plan.small_segs):
        for so, go, n in zip(segs[0::3], segs[1::3], segs[2::3]):
            name = next(nm for nm, o in store.small_off.items() if o == so)
            assert store.offsets[name] == go and store._numel[name] == n and plan.param_chunk[name] == k
            assert any(off <= go and go + n <= off + ln for off, ln in plan.ranges[k])
            seen.add(name)
    assert seen == set(store.small_off)


def test_plan_is_deterministic_and_fits_the_flag_page():
    store = ParamStore(tiny_config(), device="cpu")
    log = _synthetic_log(store, 300, seed=3)
    a = dp.plan_chunks(store, log, 300)
    b = dp.plan_chunks(store, list(log), 300)
    assert a.cuts == b.cuts and a.ranges == b.ranges and a.small_segs == b.small_segs  # identical on every rank
    assert a.n_chunks < 64


def test_staging_slot_holds_the_whole_buffer_shard():
    total = ParamStore.__new__(ParamStore)  # only the arithmetic is needed
    del total
    from sdxl_training_improvements_b200.params import unet_param_specs
    n = sum(int(np.prod(s)) for _, s in unet_param_specs(SDXL_BASE))
    assert n == 2_567_463_684
    for world in (2, 4, 8):
        _, lens, _, used = dp.shard_plan([(0, (n + 7) // 8 * 8)], world, world - 1)
        assert used <= dp.staging_slot_elems((n + 7) // 8 * 8, world)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sdxl_training_improvements_b200 import dp as D
        g = torch.zeros(64, dtype=torch.bfloat16)
        x = D.try_create_exchange(g)  # CPU tensor: the transport cannot exist; every rank must agree on "None"
        q.put((rank, x is None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_exchange_declines_collectively_without_cuda():
    import torch.multiprocessing as mp
    world = 2
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
