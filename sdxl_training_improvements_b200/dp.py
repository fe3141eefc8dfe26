"""Data-parallel gradient exchange: host side of `b2_dpx_*` (csrc/dpx.cu).

Replaces `convert_model_to_ddp` (reference: src/core/distributed.py:142-163 — DistributedDataParallel's bucketed NCCL
all-reduce that overlaps `loss.backward()`, ddpm_trainer.py:271).  One process per GPU; parameters replicated; the flat
bf16 gradient buffer of `ParamStore` is summed over ranks (the 1/world factor is folded into the optimizer's grad scale).

What is here:
  * `plan_chunks`  — from a log of which backward-tape position last writes each parameter's gradient, cut the backward
    pass into K segments and assign every parameter to the first cut after its last write: chunk k of the gradient buffer
    is final once the tape has passed `cuts[k]` and can be exchanged while the rest of the backward pass runs;
  * `shard_plan`   — this rank's shard (reduce-scatter ownership) of every contiguous piece of a chunk;
  * `PeerGradExchange` — IPC set-up (handles travel through `torch.distributed.all_gather_object`), the staging slots,
    a self-test against known sums, and `exchange_chunk(k)` / `exchange_all()` / `finish()` over the C ABI.
The pure-Python planning functions are exercised on CPU (tests/test_dp_plan.py); the transport needs >= 2 GPUs
(tests/test_gpu_dp_exchange.py).
"""
from __future__ import annotations

import bisect
import ctypes as C
import os
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

ALIGN = 8  # elements = 16 bytes: granularity of every shard (vector accesses of the reduce kernel, copy alignment)


# ---------------------------------------------------------------------------------------------------------------------
# planning (pure Python)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class ChunkPlan:
    cuts: List[int]                              # tape position (replay order) after which chunk k is final; ascending
    ranges: List[List[Tuple[int, int]]]          # chunk k -> [(element offset, length)] in the flat gradient buffer
    small_segs: List[List[int]]                  # chunk k -> flat [staging off, grad off, n, ...] triples to flush at cut k
    param_chunk: Dict[str, int] = field(default_factory=dict)
    n_tape: int = 0

    @property
    def n_chunks(self) -> int:
        return len(self.cuts)

    def describe(self) -> str:
        rows = []
        for k, (c, rg) in enumerate(zip(self.cuts, self.ranges)):
            n = sum(ln for _, ln in rg)
            rows.append(f"chunk {k}: final after tape[{c}] of {self.n_tape}, {n * 2 / 1e6:9.1f} MB in {len(rg)} piece(s), "
                        f"{len(self.small_segs[k]) // 3} small params")
        return "\n".join(rows)


def _layout(store):
    """[(offset, padded length, name)] of the flat buffer in address order; padding belongs to the preceding parameter."""
    items = sorted((off, name) for name, off in store.offsets.items())
    out = []
    for i, (off, name) in enumerate(items):
        end = items[i + 1][0] if i + 1 < len(items) else store.total
        out.append((off, end - off, name))
    return out


def last_touch_positions(store, log: Sequence[Tuple[int, str, int, int]]) -> Dict[str, int]:
    """log entries: (tape position, "flat" | "small", element offset, length) recorded by ParamStore while one backward
    pass ran.  Returns, per parameter, the last position that wrote (any part of) its gradient; -1 if never written."""
    lay = _layout(store)
    flat_offs = [o for o, _, _ in lay]
    small = sorted((so, name) for name, so in store.small_off.items())
    small_offs = [o for o, _ in small]
    last = {name: -1 for _, _, name in lay}
    for pos, kind, off, n in log:
        if n <= 0:
            continue
        if kind == "flat":
            i = max(bisect.bisect_right(flat_offs, off) - 1, 0)
            while i < len(lay) and lay[i][0] < off + n:
                last[lay[i][2]] = max(last[lay[i][2]], pos)
                i += 1
        else:
            i = max(bisect.bisect_right(small_offs, off) - 1, 0)
            while i < len(small) and small[i][0] < off + n:
                last[small[i][1]] = max(last[small[i][1]], pos)
                i += 1
    return last


def plan_chunks(store, log, n_tape: int, target_chunks: int = 10, min_elems: int = 0, tail_div: int = 128) -> ChunkPlan:
    """Greedy cuts: walk the tape in replay order, cut whenever at least total/target_chunks elements have become final
    since the previous cut; whatever is left (always including everything written by the last tape entries) is the tail
    chunk, final at the end of the backward pass."""
    lay = _layout(store)
    last = last_touch_positions(store, log)
    never = [name for _, _, name in lay if last[name] < 0]
    if never:  # a gradient the log did not see could be exchanged before it is written: refuse to plan
        raise RuntimeError(f"plan_chunks: no gradient write observed for {len(never)} parameter(s), e.g. {never[:3]}")
    total = sum(n for _, n, _ in lay)
    target = max(total // max(target_chunks, 1), min_elems, 1)
    by_pos: Dict[int, int] = {}
    for _, n, name in lay:
        by_pos[last[name]] = by_pos.get(last[name], 0) + n
    cuts: List[int] = []
    acc = 0
    for pos in sorted(by_pos):
        acc += by_pos[pos]
        if acc >= target and pos < n_tape - 1 and (not cuts or cuts[-1] < max(pos, 0)):
            cuts.append(max(pos, 0))
            acc = 0
    if not cuts or cuts[-1] != n_tape - 1:
        cuts.append(n_tape - 1)
    # the tail chunk is the only one whose exchange cannot hide behind compute: keep it small (<= total / tail_div) with
    # one extra cut as late as possible
    tail_max = max(total // max(tail_div, 1), 1)
    suffix, late = 0, None
    for pos in sorted(by_pos, reverse=True):
        if suffix + by_pos[pos] > tail_max:
            late = pos
            break
        suffix += by_pos[pos]
    if late is not None and 0 <= late < n_tape - 1 and late not in cuts and (len(cuts) < 2 or late > cuts[-2]):
        cuts.insert(len(cuts) - 1, late)
    # a parameter goes to the first chunk whose cut is at or after its last write
    chunk_of = {}
    for _, _, name in lay:
        chunk_of[name] = bisect.bisect_left(cuts, last[name])
    ranges: List[List[Tuple[int, int]]] = [[] for _ in cuts]
    for off, n, name in lay:
        r = ranges[chunk_of[name]]
        if r and r[-1][0] + r[-1][1] == off:
            r[-1] = (r[-1][0], r[-1][1] + n)
        else:
            r.append((off, n))
    segs: List[List[int]] = [[] for _ in cuts]
    for name, so in store.small_off.items():
        segs[chunk_of[name]] += [so, store.offsets[name], store._numel[name]]
    return ChunkPlan(cuts=cuts, ranges=ranges, small_segs=segs, param_chunk=chunk_of, n_tape=n_tape)


def shard_plan(ranges: Sequence[Tuple[int, int]], world: int, rank: int):
    """Reduce-scatter ownership: every piece is split into `world` shards of ceil(len / world) rounded up to 16 bytes;
    returns ([shard offset], [shard length], [offset inside a staging slot], staging elements used)."""
    offs, lens, soffs = [], [], []
    s = 0
    for off, n in ranges:
        assert off % ALIGN == 0 and n % ALIGN == 0, "pieces of the gradient buffer are 16-byte granular"
        ss = (-(-n // world) + ALIGN - 1) // ALIGN * ALIGN
        lo = min(rank * ss, n)
        hi = min(lo + ss, n)
        offs.append(off + lo)
        lens.append(hi - lo)
        soffs.append(s)
        s += ss
    return offs, lens, soffs, s


def staging_slot_elems(total: int, world: int, max_ranges: int = 4096) -> int:
    return (-(-total // world) + ALIGN * (max_ranges + 1) + ALIGN - 1) // ALIGN * ALIGN


# ---------------------------------------------------------------------------------------------------------------------
# transport
# ---------------------------------------------------------------------------------------------------------------------
MODES = ("ce_pull", "ce_push", "sm")


class PeerGradExchange:
    """Sum of the flat bf16 gradient buffer over the ranks of `group`, chunk by chunk, over peer memory (csrc/dpx.cu).
    `mode`: "ce_pull" | "ce_push" | "sm" | "auto" (time the three on a 256 MB piece at start-up, all ranks agree on the
    winner; a copy-engine transport is preferred when it is within 1.25x of the kernel one because it uses no SM)."""

    def __init__(self, grad: torch.Tensor, group=None, n_copy_streams: int = 0, self_test: bool = True,
                 mode: Optional[str] = None, autotune: bool = True):
        from . import _lib
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradExchange needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world < 2:
            raise RuntimeError("PeerGradExchange needs world_size >= 2")
        if not grad.is_cuda or grad.dtype != torch.bfloat16 or not grad.is_contiguous():
            raise RuntimeError("PeerGradExchange: the gradient buffer must be a contiguous CUDA bf16 tensor")
        if grad.numel() % ALIGN:
            raise RuntimeError("PeerGradExchange: the gradient buffer length must be a multiple of 8 elements")
        mode = (mode or os.environ.get("B2_DPX_MODE", "auto")).lower()
        if mode not in MODES + ("auto",):
            raise RuntimeError(f"PeerGradExchange: unknown mode {mode!r}")
        n_copy_streams = n_copy_streams or int(os.environ.get("B2_DPX_COPY_STREAMS", "0"))
        self.grad = grad
        self.total = grad.numel()
        self.max_chunks = int(self.lib.b2_dpx_max_chunks())
        self.WHOLE = self.max_chunks - 1   # chunk id of exchange_all()
        self.PROBE = self.max_chunks - 2   # chunk id of the start-up timing
        self.plan: Optional[ChunkPlan] = None
        self._chunk_args: Dict[int, tuple] = {}
        self._seg_dev: Dict[int, torch.Tensor] = {}
        self.seq = 1
        self.issued = False
        self.handles: Dict[str, C.c_void_p] = {}
        self.timings: Dict[str, float] = {}
        with torch.cuda.device(grad.device):
            self.slot = staging_slot_elems(self.total, self.world)
            self.staging = torch.empty((self.world - 1) * self.slot, device=grad.device, dtype=torch.bfloat16)
            flags = C.c_void_p()
            _lib.check(self.lib.b2_dpx_alloc_flags(C.byref(flags)), "dpx_alloc_flags")

            def export(ptr):
                h = (C.c_ubyte * 64)()
                off = C.c_int64()
                _lib.check(self.lib.b2_dpx_ipc_export(C.c_void_p(ptr), h, C.byref(off)), "dpx_ipc_export")
                return bytes(h), int(off.value)

            mine = (export(grad.data_ptr()), export(flags.value), export(self.staging.data_ptr()), self.total)
            everyone: List = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)

            def imp(hb, off):
                ptr = C.c_void_p()
                _lib.check(self.lib.b2_dpx_ipc_import((C.c_ubyte * 64).from_buffer_copy(hb), off, C.byref(ptr)),
                           "dpx_ipc_import")
                return ptr.value

            gp = (C.c_void_p * self.world)()
            fp = (C.c_void_p * self.world)()
            sp = (C.c_void_p * self.world)()
            for p, (g_h, f_h, s_h, ptotal) in enumerate(everyone):
                if ptotal != self.total:
                    raise RuntimeError("PeerGradExchange: ranks disagree on the gradient buffer size")
                if p == self.rank:
                    gp[p], fp[p], sp[p] = grad.data_ptr(), flags.value, self.staging.data_ptr()
                else:
                    gp[p], fp[p], sp[p] = imp(*g_h), imp(*f_h), imp(*s_h)
            for m in (MODES if mode == "auto" else (mode,)):
                if m == "sm" and self.world > 8:
                    continue
                h = C.c_void_p()
                _lib.check(self.lib.b2_dpx_create(self.rank, self.world, MODES.index(m), gp, fp, sp, self.slot,
                                                  int(n_copy_streams), C.byref(h)), f"dpx_create({m})")
                self.handles[m] = h
        # before (or without) the start-up timing: the transport that won every measurement so far
        self.mode = "ce_push" if "ce_push" in self.handles else next(iter(self.handles))
        self._set_chunk(self.WHOLE, [(0, self.total)], 0)
        if self_test:
            self.self_test()
        if autotune and len(self.handles) > 1:
            self.autotune()

    @property
    def handle(self):
        return self.handles[self.mode]

    # ---- plan ----
    def _set_chunk(self, k: int, ranges, staging_base: int = 0):
        _, _, _, used = shard_plan(ranges, self.world, self.rank)
        assert staging_base % ALIGN == 0 and staging_base + used <= self.slot, "staging slot too small for this chunk"
        n = len(ranges)
        arr = lambda v: (C.c_int64 * max(n, 1))(*v)  # noqa: E731
        self._chunk_args[k] = (n, arr([o for o, _ in ranges]), arr([ln for _, ln in ranges]), int(staging_base))
        return staging_base + used

    def set_plan(self, plan: ChunkPlan):
        if plan.n_chunks > self.max_chunks - 2:
            raise RuntimeError(f"PeerGradExchange: {plan.n_chunks} chunks > {self.max_chunks - 2}")
        self.plan = plan
        self._seg_dev = {}
        base = 0  # every chunk gets its own region of the staging slots (the push transport relies on it)
        for k, rg in enumerate(plan.ranges):
            base = self._set_chunk(k, rg, base)
            if plan.small_segs[k]:
                self._seg_dev[k] = torch.tensor(plan.small_segs[k], dtype=torch.int64, device=self.grad.device)

    def flush_chunk(self, store, k: int):
        """Fold the fp32-staged small-parameter gradients that belong to chunk k into the flat buffer (capturable)."""
        from . import ops
        if k == self.plan.n_chunks - 1:
            store.flush_small_grads()  # tail: everything that is left
        elif k in self._seg_dev:
            seg = self._seg_dev[k]
            ops.flush_small_grads(store.small32, store.grad, seg, seg.numel() // 3)

    # ---- transport ----
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.grad.device).cuda_stream)

    def exchange_chunk(self, k: int):
        from . import _lib
        n, offs, lens, base = self._chunk_args[k]
        _lib.check(self.lib.b2_dpx_exchange(self.handle, int(k), C.c_uint32(self.seq & 0xFFFFFFFF), n, offs, lens, base,
                                            self._stream()), f"dpx_exchange({self.mode})")
        self.issued = True

    def exchange_all(self):
        """The whole buffer as one chunk (no overlap): the first step before a plan exists, or a caller that did not
        announce the last accumulation micro-step."""
        self.exchange_chunk(self.WHOLE)

    @property
    def provides_norm(self) -> bool:
        """True when finish(gnorm_sq_out=...) can return the squared norm of the reduced gradients (copy-engine transports)."""
        return bool(self.lib.b2_dpx_norm_supported(self.handle)) and os.environ.get("B2_DP_NORM_FOLD", "1") != "0"

    def finish(self, gnorm_sq_out: Optional[torch.Tensor] = None):
        """The current stream waits for every chunk of this step to be complete in the local buffer.  With `gnorm_sq_out`
        (a float64 device scalar) it also receives sum((reduced gradient)^2) over the whole buffer, accumulated by the
        exchange's reduce kernels and added over ranks in rank order — the optimizer's clip needs no pass of its own."""
        from . import _lib
        if gnorm_sq_out is not None:
            assert gnorm_sq_out.dtype == torch.float64 and gnorm_sq_out.is_cuda
            _lib.check(self.lib.b2_dpx_finish_norm(self.handle, C.c_uint32(self.seq & 0xFFFFFFFF), self._stream(),
                                                   C.c_void_p(gnorm_sq_out.data_ptr())), "dpx_finish_norm")
        else:
            _lib.check(self.lib.b2_dpx_finish(self.handle, C.c_uint32(self.seq & 0xFFFFFFFF), self._stream()), "dpx_finish")
        self.seq += 1
        self.issued = False

    def _wait_or_die(self, what: str, timeout_s: float = 60.0):
        """A stuck flag would block the stream for ever: poll for completion and exit instead of hanging the job."""
        ev = torch.cuda.Event()
        ev.record()
        t0 = time.time()
        while not ev.query():
            if time.time() - t0 > timeout_s:
                print(f"[dpx] rank {self.rank}: {what} did not complete in {timeout_s:.0f} s — aborting", flush=True)
                os._exit(17)
            time.sleep(0.002)

    def self_test(self):
        """Known-answer exchange of the (zero) gradient buffer with every transport: rank r fills element i with
        (r + 1) * (i % 5 + 1), small integers whose sums are exact in bf16."""
        g = self.grad
        idx = torch.arange(self.total, device=g.device, dtype=torch.int32) % 5 + 1
        want = (idx * (self.world * (self.world + 1) // 2)).to(torch.bfloat16)
        keep = self.mode
        for m in self.handles:
            self.mode = m
            g.copy_((idx * (self.rank + 1)).to(torch.bfloat16))
            torch.cuda.synchronize(g.device)
            torch.distributed.barrier(group=self.group)
            self.exchange_all()
            self.finish()
            self._wait_or_die(f"self-test exchange ({m})")
            bad = int((g != want).sum())
            g.zero_()
            torch.cuda.synchronize(g.device)
            torch.distributed.barrier(group=self.group)
            if bad:
                self.mode = keep
                raise RuntimeError(f"PeerGradExchange self-test ({m}): {bad} of {self.total} elements differ from the "
                                   "known sums")
        self.mode = keep

    def autotune(self, probe_elems: int = 128 * 1024 * 1024, iters: int = 3):
        """Time every transport on one piece of the (idle, zero) buffer; all ranks pick the same winner."""
        dist = torch.distributed
        n = min(self.total, probe_elems) // ALIGN * ALIGN
        self._set_chunk(self.PROBE, [(0, n)], 0)
        names = list(self.handles)
        t = torch.zeros(len(names), device=self.grad.device)
        for i, m in enumerate(names):
            self.mode = m
            for it in range(iters + 1):
                if it == 1:
                    torch.cuda.synchronize(self.grad.device)
                    dist.barrier(group=self.group)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                self.exchange_chunk(self.PROBE)
                self.finish()
            e1.record()
            self._wait_or_die(f"autotune ({m})")
            t[i] = e0.elapsed_time(e1) / iters
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.timings = {m: float(t[i]) for i, m in enumerate(names)}
        best = min(names, key=lambda m: self.timings[m])
        ce = [m for m in names if m != "sm"]
        if best == "sm" and ce:
            alt = min(ce, key=lambda m: self.timings[m])
            if self.timings[alt] <= 1.25 * self.timings["sm"]:
                best = alt
        self.mode = best
        self.probe_gbs = {m: 2 * (self.world - 1) / self.world * n * 2 / (ms * 1e-3) / 1e9 for m, ms in self.timings.items()}
        if self.rank == 0:
            print("[dpx] transports (ms for %.0f MB, NVLink ingress GB/s per GPU): %s -> %s" % (
                n * 2 / 1e6, ", ".join(f"{m} {self.timings[m]:.2f} ms / {self.probe_gbs[m]:.0f}" for m in names), best),
                flush=True)
        self.grad.zero_()
        torch.cuda.synchronize(self.grad.device)

    def close(self):
        for h in self.handles.values():
            self.lib.b2_dpx_destroy(h)
        self.handles = {}


def try_create_exchange(grad: torch.Tensor, group=None) -> Optional[PeerGradExchange]:
    """Collective: every rank either gets an exchange object or None (then the caller uses one NCCL all-reduce).
    `B2_DP_EXCHANGE=nccl` forces the NCCL path; `B2_DPX_MODE` picks the transport (default auto)."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return None
    if os.environ.get("B2_DP_EXCHANGE", "peer").lower() == "nccl":
        return None
    x, err = None, ""
    try:
        x = PeerGradExchange(grad, group=group, self_test=False, autotune=False)
    except Exception as e:  # noqa: BLE001
        err = f"{type(e).__name__}: {e}"
    ok = torch.tensor([1 if x is not None else 0], device=grad.device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok) == 1:
        try:
            x.self_test()
        except Exception as e:  # noqa: BLE001
            err = f"{type(e).__name__}: {e}"
            x = None
        ok = torch.tensor([1 if x is not None else 0], device=grad.device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok) != 1:
        if dist.get_rank(group) == 0 or err:
            print(f"[dpx] peer-memory gradient exchange unavailable ({err or 'another rank failed'}); using NCCL all-reduce",
                  flush=True)
        return None
    if len(x.handles) > 1:
        x.autotune()
    return x
