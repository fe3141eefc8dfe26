#!/usr/bin/env python
"""bench.py — SDXL-base 1024^2 bf16 training images/sec on N x B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--method ddpm|flow_matching]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one optimizer step of the reference's default micro-batch (bs=4 per GPU, latent [4,4,128,128], text
[4,77,2048]): Philox noise + noising + full SDXL UNet forward + full backward + MSE loss (+ one all-reduce of the flat
gradient buffer when N>1) + grad-norm clip + fused AdamW.  Synthetic data, seeded random-init weights (no checkpoints
offline).  `value` is timed with the batch resident in HBM; `e2e` goes through the plugin API
(`B200DDPMTrainer._execute_training_step(batch)`) with pinned-host inputs copied H2D and the loss/metrics read back
D2H inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "sdxl_base_1024px_bf16_train_images_per_sec"
# one workload string for both arms (the driver compares `config` across `--impl ours` / `--impl reference`)
WORKLOAD_FMT = ("SDXL-base UNet, {method} v_prediction, bs=4/GPU, {W}x{H} (latent {h}x{w}), {accum}bf16, "
                "full fwd+bwd+loss+clip+{opt} (configs[1])")
IMG_FLOPS_1024 = None  # filled from the analytical model


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=float(d["bf16_tflops"]), sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    hbm=float(d["hbm_gbs"]), src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


def _config_ns(method: str):
    return SimpleNamespace(
        model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0, use_ztsnr=True,
                              min_snr_gamma=None, prediction_type="v_prediction"),
        training=SimpleNamespace(method=method, prediction_type="v_prediction", gradient_accumulation_steps=1,
                                 clip_grad_norm=1.0, batch_size=4))


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def stop(self):
        self._halt.set()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------
def _init_weights_(unet, seed: int):
    """Seeded synthetic init on device: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights, unit norm scales, small biases."""
    g = torch.Generator(device=unet.device).manual_seed(seed)
    with torch.no_grad():
        for name, p in unet.named_parameters():
            if p.dim() >= 2:
                bound = 1.0 / (p[0].numel() ** 0.5)
                p.copy_((torch.rand(p.shape, device=p.device, generator=g) * 2 - 1) * bound)
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.copy_((torch.rand(p.shape, device=p.device, generator=g) * 2 - 1) * 0.02)


def _synthetic_batch(B, H, W, seed, pin=True):
    g = torch.Generator().manual_seed(seed)
    b = {
        "vae_latents": torch.randn(B, 4, H, W, generator=g),
        "prompt_embeds": torch.randn(B, 77, 2048, generator=g).to(torch.bfloat16),
        "pooled_prompt_embeds": torch.randn(B, 1280, generator=g).to(torch.bfloat16),
        "time_ids": torch.tensor([[8.0 * W, 8.0 * H, 0, 0, 8.0 * W, 8.0 * H]]).repeat(B, 1)[:, None],
    }
    if pin:
        b = {k: v.pin_memory() for k, v in b.items()}
    b["metadata"] = [{} for _ in range(B)]
    return b


def _time_gemm_roofline(ops, peaks):
    """Dominant kernel = gemm2_kernel (persistent CTA-pair tcgen05 GEMM, ~50 % of the step); timed live on the GEGLU
    up-projection shape (the FFN is 41.7 % of the step's FLOPs)."""
    M, N, K = 4096, 10240, 1280
    nset = 6  # rotate operands: 6 x (26 MB W + 84 MB D) >> 126 MB L2
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    Ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) * 0.03 for _ in range(nset)]
    Ds = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(nset)]
    for i in range(nset):
        ops.linear_fwd(x, Ws[i], out=Ds[i])
    torch.cuda.synchronize()
    iters = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.linear_fwd(x, Ws[i % nset], out=Ds[i % nset])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * M * N * K
    ach = flops / (ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    agg = None  # all gemm2 launches of one training step, from the committed ncu launch list joined with the shape log
    ap = os.path.join(ROOT, "profiles", "gemm_aggregate.json")
    if os.path.exists(ap):
        try:
            agg = json.load(open(ap))
            agg["frac_of_burst_peak"] = round(agg["tflops"] / peaks["burst"], 4)
        except Exception:
            agg = None
    return {"bound": "tensor", "achieved": round(ach, 1), "peak": peaks["burst"], "unit": "TFLOP/s",
            "frac": round(ach / peaks["burst"], 4), "traffic": traffic, "aggregate_all_gemm_launches": agg,
            "kernel": "gemm2_kernel (cta_group::2, 256xBN tiles) on M=4096,N=10240,K=1280 (GEGLU up-projection), "
                      f"{flops / 1e9:.1f} GFLOP/launch, {ms * 1e3:.1f} us/launch, peak = {peaks['src']} burst bf16"}


def _graph_time_us(fn, iters=10):
    """Average device time of `fn` (a few kernel launches) inside a CUDA-graph replay of `iters` repetitions."""
    fn()
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def _extra_rooflines(ops, peaks):
    """BASELINE.json asks for the attention / conv fraction of roofline next to images/s: the other hot kernels timed
    live (CUDA-graph replay, CUDA events) at their SDXL shapes (B=4, 1024^2).  Peak = measured burst bf16 / measured
    copy bandwidth (MEASURED_PEAKS.json).  FLOPs are algorithmic: attention core 4 n_q n_k 64 per head
    forward, 10 n_q n_k 64 backward; conv 2 M Cout 9 Cin; the cross-attention "fused block" is LN -> Wq -> softmax(QK^T)V over
    77 keys -> Wo + bias + residual = 28.46 GFLOP per layer call at C=1280 (the north-star figure; four launches today)."""
    bf = torch.bfloat16
    out = {}
    try:
        for tag, (B, H, n) in (("self_attn_n4096", (4, 10, 4096)), ("self_attn_n1024", (4, 20, 1024))):
            Cc = H * 64
            qkv = torch.randn(B * n, 3 * Cc, device="cuda").to(bf)
            q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
            do = torch.randn(B * n, Cc, device="cuda").to(bf)
            dqkv = torch.empty_like(qkv)
            o, lse = ops.attn_fwd(q, k, v, B, H, n, n, 0.125)
            tf = _graph_time_us(lambda: ops.attn_fwd(q, k, v, B, H, n, n, 0.125, out=o))
            tb = _graph_time_us(lambda: ops.attn_bwd(q, k, v, o, lse, do, dqkv[:, :Cc], dqkv[:, Cc:2 * Cc], dqkv[:, 2 * Cc:],
                                                     B, H, n, n, 0.125))
            ff = 4.0 * n * n * 64 * B * H
            out[tag] = {"fwd_us": round(tf, 1), "fwd_tflops": round(ff / tf / 1e6, 1),
                        "fwd_frac_of_peak": round(ff / tf / 1e6 / peaks["burst"], 3),
                        "bwd_us": round(tb, 1), "bwd_tflops": round(2.5 * ff / tb / 1e6, 1),
                        "bwd_frac_of_peak": round(2.5 * ff / tb / 1e6 / peaks["burst"], 3), "bound": "d=64: ex2 (16 384 per 128x128 block = 1 024 XU cycles) and the S MMA's shared-memory operand reads (128 B/clk) both sit at ~1 000 cycles per block; measured block period ~1 450-1 600 (hand-off latency between softmax groups and the MMA thread) - profiles/r2_attention_notes.md"}
        # cross-attention, C=1280, n=1024, 77 keys: the core alone (HBM-bound) and the fused-block definition
        B, H, n, nk = 4, 20, 1024, 77
        Cc = H * 64
        x = torch.randn(B * n, Cc, device="cuda").to(bf)
        gamma, beta = torch.ones(Cc, device="cuda", dtype=bf), torch.zeros(Cc, device="cuda", dtype=bf)
        Wq = (torch.randn(Cc, Cc, device="cuda") * 0.02).to(bf)
        Wo = (torch.randn(Cc, Cc, device="cuda") * 0.02).to(bf)
        bo = torch.zeros(Cc, device="cuda", dtype=bf)
        kv = torch.randn(B * nk, 2 * Cc, device="cuda").to(bf)
        kk, vv = kv[:, :Cc], kv[:, Cc:]
        qb = torch.randn_like(x)
        ob = torch.empty_like(x)
        yb = torch.empty_like(x)
        tcore = _graph_time_us(lambda: ops.attn_fwd(qb, kk, vv, B, H, n, nk, 0.125, out=ob))

        fused = bool(ops.xattn_q_core_ok(B, n, nk, Cc))

        def block():  # what UNetEngine.attention() launches for one cross-attention layer call
            xn, _, _ = ops.ln_fwd(x, gamma, beta, 1e-5)
            if fused:  # to_q GEMM + 77-key core in one launch (csrc/xattn.cu)
                _, o_, _ = ops.xattn_q_core(xn, Wq, kk, vv, B, n, nk, 0.125)
            else:
                ops.linear_fwd(xn, Wq, out=qb)
                o_ = ops.attn_fwd(qb, kk, vv, B, H, n, nk, 0.125, out=ob)[0]
            ops.linear_fwd(o_, Wo, bias=bo, residual=x, out=yb)

        tblock = _graph_time_us(block)
        tqcore = _graph_time_us(lambda: ops.xattn_q_core(x, Wq, kk, vv, B, n, nk, 0.125)) if fused else None
        # backward of the core (dQ, dK, dV from dO): one pass, P / dS tiles in shared memory (csrc/xattn_bwd.cu)
        ob2, lse2 = ops.attn_fwd(qb, kk, vv, B, H, n, nk, 0.125)
        dob = torch.randn_like(x)
        dqb, dkvb = torch.empty_like(x), torch.empty_like(kv)
        tcore_bwd = _graph_time_us(lambda: ops.attn_bwd(qb, kk, vv, ob2, lse2, dob, dqb, dkvb[:, :Cc], dkvb[:, Cc:],
                                                        B, H, n, nk, 0.125))
        core_flops = 4.0 * n * nk * 64 * B * H
        core_bytes = 2 * x.numel() * 2 + kv.numel() * 2
        block_flops = 2 * (2.0 * B * n * Cc * Cc) + core_flops
        out["cross_attn_c1280"] = {
            "core_us": round(tcore, 1), "core_gbs": round(core_bytes / tcore / 1e3, 0),
            "core_frac_of_hbm_peak": round(core_bytes / tcore / 1e3 / peaks["hbm"], 3),
            "core_tflops": round(core_flops / tcore / 1e6, 1),
            "fused_block_us": round(tblock, 1), "fused_block_gflop": round(block_flops / 1e9, 2),
            "fused_block_tflops": round(block_flops / tblock / 1e6, 1),
            "fused_block_frac_of_peak": round(block_flops / tblock / 1e6 / peaks["burst"], 3),
            "north_star_target_frac": 0.6, "launches_per_block": 3 if fused else 4,
            "q_proj_plus_core_us": round(tqcore, 1) if tqcore else None,
            "q_proj_plus_core_tflops": round((2.0 * B * n * Cc * Cc + core_flops) / tqcore / 1e6, 1) if tqcore else None,
            "core_bwd_us": round(tcore_bwd, 1)}
        # GEGLU up-projection (ff1), M=4096, 2F=10240, K=1280: gate in the GEMM epilogue vs GEMM + gate kernel
        Mg, Fg, Kg = 4096, 5120, 1280
        xg = torch.randn(Mg, Kg, device="cuda").to(bf)
        W1 = (torch.randn(2 * Fg, Kg, device="cuda") * 0.03).to(bf)
        b1 = torch.zeros(2 * Fg, device="cuda", dtype=bf)
        ug = torch.empty(Mg, 2 * Fg, device="cuda", dtype=bf)
        gg = {"gemm_gflop": round(2.0 * Mg * 2 * Fg * Kg / 1e9, 1)}
        if ops.linear_geglu_ok(Mg, Fg, Kg):
            gg["fused_us"] = round(_graph_time_us(lambda: ops.linear_geglu_fwd(xg, W1, b1, Fg)), 1)
        gg["gemm_us"] = round(_graph_time_us(lambda: ops.linear_fwd(xg, W1, bias=b1, out=ug)), 1)
        gg["gate_kernel_us"] = round(_graph_time_us(lambda: ops.geglu_fwd(ug, Fg)), 1)
        out["geglu_ff1"] = gg
        # implicit-GEMM 3x3 conv 640 -> 640 @ 64^2, B=4
        B, Hh, Ww, Ci, Co = 4, 64, 64, 640, 640
        M = B * Hh * Ww
        xc = torch.randn(M, Ci, device="cuda").to(bf)
        wk = (torch.randn(Co, 9 * Ci, device="cuda") / (9 * Ci) ** 0.5).to(bf)
        dy = torch.randn(M, Co, device="cuda").to(bf)
        dx = torch.empty_like(xc)
        dw = torch.zeros_like(wk)
        yc = torch.empty(M, Co, device="cuda", dtype=bf)
        bias = torch.zeros(Co, device="cuda", dtype=bf)
        cf = 2.0 * M * Co * 9 * Ci
        t1 = _graph_time_us(lambda: ops.conv3x3_fwd(xc, wk, B, Hh, Ww, Ci, Co, bias=bias, out=yc))
        t2 = _graph_time_us(lambda: ops.conv3x3_dgrad(dy, wk, dx, B, Hh, Ww, Ci, Co))
        t3 = _graph_time_us(lambda: ops.conv3x3_wgrad(dy, xc, dw, B, Hh, Ww, Ci, Co))
        out["conv3x3_640_64x64"] = {k: v for k, v in (
            ("gflop", round(cf / 1e9, 1)),
            ("fwd_us", round(t1, 1)), ("fwd_frac_of_peak", round(cf / t1 / 1e6 / peaks["burst"], 3)),
            ("dgrad_us", round(t2, 1)), ("dgrad_frac_of_peak", round(cf / t2 / 1e6 / peaks["burst"], 3)),
            ("wgrad_us", round(t3, 1)), ("wgrad_frac_of_peak", round(cf / t3 / 1e6 / peaks["burst"], 3)))}
    except Exception as e:  # noqa: BLE001  (diagnostic extras must never lose the main measurement)
        out["error"] = f"{type(e).__name__}: {str(e)[:160]}"
    return out


CPU_SAMPLE_STEP_BUDGET_S = 9.0   # per CPU step; both CPU legs use the SAME rule, so on one box they time the same sample


def _cpu_baseline(threads=None, timed_steps=1):
    """The oracle (PyTorch-eager bf16 on the host CPU = the reference's own CPU path) on a bounded sample of configs[1]:
    fwd + bwd + MSE of the full SDXL UNet on ONE image (B=1) at the largest latent of (128, 96, 64, 48, 32) whose step
    fits CPU_SAMPLE_STEP_BUDGET_S, scaled to 1024^2 images by algorithmic FLOPs.  Used by BOTH `--impl reference` and the
    `cpu_baseline` leg of the main arm (same rule -> same sample on the same box).  Conservative choices, stated: bf16
    (AMX) is ~2.7x faster on these hosts than the fp32 eager SURVEY 8d names, and no optimizer step is timed."""
    from oracle.unet_sdxl import OracleUNet
    from sdxl_training_improvements_b200.flops import train_step_flops
    if threads is None:  # every host core this process may use (torchrun exports OMP_NUM_THREADS=1)
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    if torch.get_num_threads() != threads:
        torch.set_num_threads(threads)
    t0 = time.time()
    with torch.device("meta"):
        m = OracleUNet()
    m = m.to(torch.bfloat16).to_empty(device="cpu")
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() >= 2:
                p.uniform_(-0.02, 0.02)
            else:
                p.fill_(0.5)
    build_s = time.time() - t0

    def one(hw):
        x = torch.randn(1, 4, hw, hw).to(torch.bfloat16)
        ctx = torch.randn(1, 77, 2048).to(torch.bfloat16)
        pooled = torch.randn(1, 1280).to(torch.bfloat16)
        tid = torch.tensor([[8.0 * hw, 8.0 * hw, 0, 0, 8.0 * hw, 8.0 * hw]])
        t = time.time()
        out = m(x, torch.tensor([500]), ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
        out.float().square().mean().backward()
        for p in m.parameters():
            p.grad = None
        return time.time() - t

    probe = one(32)  # also warms the allocator / threads
    probe = min(probe, one(32))
    full = train_step_flops(None, 128, 128)
    best = 32
    for hw in (128, 96, 64, 48):
        est = probe * train_step_flops(None, hw, hw) / train_step_flops(None, 32, 32)
        if est <= CPU_SAMPLE_STEP_BUDGET_S:
            best = hw
            break
    dts = [one(best) for _ in range(max(1, timed_steps))]
    dt = sum(dts) / len(dts)
    frac = train_step_flops(None, best, best) / full
    return m, {"value": round(frac / dt, 5), "unit": "images/s", "cores": threads, "kind": "port",
               "sample": f"oracle (PyTorch eager bf16, CPU) fwd+bwd+MSE of the full SDXL UNet, B=1, latent {best}x{best} "
                         f"({dt:.1f} s/step), scaled to 1024^2-image equivalents by algorithmic FLOPs ({frac:.3f} img/step); "
                         f"no optimizer step; model build {build_s:.0f} s untimed",
               "sample_latent": best, "precision": "bf16 (AMX) — faster than fp32 eager on this host: conservative"}, one


def _torch_eager_gpu_baseline(B, H, W, steps=3):
    """The reference's own GPU path on this box: the oracle module tree (= diffusers' UNet2DConditionModel restated) in
    PyTorch eager bf16 on cuda:0 -> cuDNN convs, cuBLASLt linears, SDPA attention, autograd backward, + MSE loss.  No
    optimizer step is timed here (the reference's AdamWBF16 is ~20 eager kernels per tensor), so this flatters it."""
    from oracle.unet_sdxl import OracleUNet
    try:
        with torch.device("meta"):
            m = OracleUNet()
        m = m.to(torch.bfloat16).to_empty(device="cuda")
        with torch.no_grad():
            for p in m.parameters():
                if p.dim() >= 2:
                    p.uniform_(-0.02, 0.02)
                else:
                    p.fill_(0.5)
        x = torch.randn(B, 4, H, W, device="cuda", dtype=torch.bfloat16)
        ctx = torch.randn(B, 77, 2048, device="cuda", dtype=torch.bfloat16)
        pooled = torch.randn(B, 1280, device="cuda", dtype=torch.bfloat16)
        tid = torch.tensor([[8.0 * W, 8.0 * H, 0, 0, 8.0 * W, 8.0 * H]], device="cuda").repeat(B, 1)
        t = torch.randint(0, 1000, (B,), device="cuda")
        tgt = torch.randn(B, 4, H, W, device="cuda", dtype=torch.bfloat16)

        def step():
            out = m(x, t, ctx, added_cond_kwargs={"text_embeds": pooled, "time_ids": tid}).sample
            torch.nn.functional.mse_loss(out, tgt).backward()

        def best_of(fn):
            fn()
            fn()  # two warm-ups: cuDNN / cuBLASLt heuristics and the caching allocator settle on the second pass
            torch.cuda.synchronize()
            best = None
            for _ in range(max(2, steps)):  # best of N: this is a baseline, give it every benefit
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                t_ = e0.elapsed_time(e1)
                best = t_ if best is None else min(best, t_)
            return best

        ms = best_of(step)
        res = {"value": round(B / (ms * 1e-3), 3), "unit": "images/s", "ms_per_step": round(ms, 1),
               "what": f"oracle UNet (diffusers module tree) in PyTorch eager bf16 on the same GPU, fwd+bwd+MSE, B={B}, "
                       "no optimizer step, best of 3 after 2 warm-ups"}
        try:  # the same + global-norm clip + torch's FUSED AdamW (the library's best case; the reference's own AdamWBF16
            # is ~20 eager kernels per tensor x 1,680 tensors and would be far slower)
            opt = torch.optim.AdamW(m.parameters(), lr=4e-7, weight_decay=1e-2, fused=True)

            def full_step():
                step()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0, foreach=True)
                opt.step()
                opt.zero_grad(set_to_none=True)

            ms2 = best_of(full_step)
            res["with_optimizer"] = {"value": round(B / (ms2 * 1e-3), 3), "ms_per_step": round(ms2, 1),
                                     "what": "same + clip_grad_norm_(foreach) + torch.optim.AdamW(fused=True) — the same work "
                                             "as this repo's step"}
            opt = None
        except Exception as e:  # noqa: BLE001
            res["with_optimizer"] = {"unavailable": f"{type(e).__name__}: {str(e)[:100]}"}
    except Exception as e:  # noqa: BLE001  (an OOM here must not lose the main measurement)
        res = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    finally:
        m = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sdxl_training_improvements_b200.flops import train_step_flops
    # torchrun exports OMP_NUM_THREADS=1 to every rank; _cpu_baseline() claims all host cores this process may use
    m, cb, one = _cpu_baseline()
    hw = cb["sample_latent"]
    for _ in range(args.warmup):
        one(hw)
    t0 = time.time()
    for _ in range(args.steps):
        one(hw)
    dt = (time.time() - t0) / max(1, args.steps)
    frac = train_step_flops(None, hw, hw) / train_step_flops(None, 128, 128)
    val = frac / dt
    cb["value"] = round(val, 5)
    out = {"impl": "reference", "metric": METRIC, "value": round(val, 5), "unit": "images/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": WORKLOAD_FMT.format(method=args.method, W=8 * args.latent_w, H=8 * args.latent_h,
                                                      h=args.latent_h, w=args.latent_w, accum="", opt=args.optimizer),
                      "cpu_sample": f"B=1 latent {hw}x{hw} fwd+bwd+MSE per step, FLOP-scaled to 1024^2 images "
                                    "(same rule as the main arm's cpu_baseline leg)"},
           "cpu_baseline": cb,
           "e2e": {"value": round(val, 5), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_ours(args):
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the sm_100a kernels have no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = torch.distributed
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sdxl_training_improvements_b200 import _lib, ops
    from sdxl_training_improvements_b200.flops import train_step_flops
    from sdxl_training_improvements_b200.trainer import B200AdamW, B200AdamWBF16, create_trainer
    from sdxl_training_improvements_b200.unet import B200UNet

    peaks = _peaks()
    B, H, W = 4, args.latent_h, args.latent_w
    A = max(1, args.accum)  # micro-steps per optimizer step (config 5: grad-accum 4)
    cfg = _config_ns(args.method)
    cfg.training.gradient_accumulation_steps = A
    eager = None
    if world == 1 and not args.no_eager_baseline:
        eager = _torch_eager_gpu_baseline(B, H, W)
    unet = B200UNet(device=f"cuda:{local}")
    _init_weights_(unet, seed=1234)  # identical on every rank (replicated parameters)
    if args.optimizer == "adamw_bf16":  # the reference's default (src/config.yaml: optimizer_type adamw_bf16)
        opt = B200AdamWBF16(unet, lr=4e-7, weight_decay=1e-2, seed=4321)
    else:
        opt = B200AdamW(unet, lr=4e-7, weight_decay=1e-2, master_weights=True)
    use_graph = not args.no_graph
    trainer = create_trainer(cfg, unet, opt, device=f"cuda:{local}", seed=1000 + rank, cuda_graph=use_graph)
    if args.no_grad_exchange and world > 1:
        trainer.core.dp = None
        trainer.world_size = 1
    # BASELINE configs[4] (--mixed-buckets): aspect buckets {768^2, 1024^2, 1280x960}, the bucket of every GLOBAL micro-step
    # drawn by data.BucketBatchSampler (rank-sharded, identical order on every rank: all ranks run the same shape)
    mixed = bool(args.mixed_buckets)
    shapes = [(96, 96), (128, 128), (120, 160)] if mixed else [(H, W)]
    batches = {s_: _synthetic_batch(B, s_[0], s_[1], seed=77 + rank + 13 * i_) for i_, s_ in enumerate(shapes)}
    batch = batches[shapes[-1] if not mixed else (128, 128)]
    h2d = sum(v.numel() * v.element_size() for k, v in batch.items() if torch.is_tensor(v))
    n_micro = (args.steps + max(args.warmup, 3) + 4) * A
    if mixed:
        from sdxl_training_improvements_b200.data import BucketBatchSampler
        per_bucket = B * world * ((n_micro + len(shapes) - 1) // len(shapes) + 2)
        idx, shape_of = {}, {}
        for i_, s_ in enumerate(shapes):
            idx[s_] = list(range(i_ * per_bucket, (i_ + 1) * per_bucket))
            shape_of.update({j_: s_ for j_ in idx[s_]})
        sampler_b = BucketBatchSampler(idx, B, rank=rank, world_size=world, seed=11)
        seq = [shape_of[b_[0]] for b_ in sampler_b][:n_micro]
        assert len(seq) == n_micro
    else:
        seq = [shapes[0]] * n_micro
    seq_pos = [0]

    def next_shape():
        s_ = seq[seq_pos[0] % len(seq)]
        seq_pos[0] += 1
        return s_

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public plugin API (also JITs nothing: all kernels are prebuilt) ----
    def api_step(fixed=None):
        for a in range(A):
            s_ = fixed or next_shape()
            loss, metrics = trainer._execute_training_step(batches[s_], accumulate=A > 1, is_last_accumulation_step=a == A - 1)
        return loss, metrics

    if mixed:  # every bucket shape is first seen at an optimizer-step boundary, so its CUDA graph is captured there
        for s_ in shapes:
            api_step(fixed=s_)
    for _ in range(max(args.warmup, 3)):
        api_step()
    barrier()

    # ---- set-up shared by both timed loops ----
    from sdxl_training_improvements_b200.trainer import _prep_batch, allreduce_gradients
    dev_batches = {s_: _prep_batch(batches[s_], unet.device) for s_ in shapes}
    K = args.steps
    sched = trainer.noise_scheduler if args.method == "ddpm" else None
    gen = torch.Generator().manual_seed(5 + rank)
    if args.method == "ddpm":
        ts = [sched.sample_timesteps(B, generator=gen) for _ in range(K * A)]
        t_embed = [t.float().cuda() for t in ts]
        sig = [sched.timestep_to_sigma(t).float().cuda() for t in ts]
    else:
        from sdxl_training_improvements_b200.trainer import sample_logit_normal
        ts = [sample_logit_normal((B,), generator=gen) for _ in range(K * A)]
        t_embed = [t.float().cuda() for t in ts]
        sig = t_embed
    core = trainer.core
    gms = {s_: trainer._micro_graphs[(B, s_[0], s_[1], 77)] for s_ in shapes} if use_graph else {}
    gm = gms[shapes[0]] if use_graph else None
    og = trainer._opt_graph if use_graph else None
    timed_pos0 = seq_pos[0]
    # ---- (1) end-to-end through the plugin API with host buffers: `e2e` (the headline) ----
    # Both loops see the same timestep / bucket sequence: under the power cap a step whose loss hits the reference's clamp (zero
    # gradients through the backward GEMMs) draws less power and runs at higher clocks.  The end-to-end loop runs FIRST: on
    # these boxes whichever loop runs second is ~1 % slower (SM clock drifts down as the package heats up; tools/e2e_gap.py
    # shows < 0.15 ms of GPU idle time per step between the graphs of the API path).
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    torch.manual_seed(5 + rank)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last_loss = None
    for i in range(K):
        loss, metrics = api_step()
        last_loss = metrics["loss"]  # python float: the D2H read of the step's result already happened
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3) / K

    # ---- (2) device-resident timing: `value` ----
    if mixed:
        seq_pos[0] = timed_pos0
    timed_seq = [next_shape() for _ in range(K * A)]
    launches0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        for a in range(A):
            core.dp_last = a == A - 1  # N>1: the last micro-step's backward hands finished gradient chunks to the exchange
            s_ = timed_seq[i * A + a]
            dev_batch = dev_batches[s_]
            if use_graph:  # one graph launch per micro-step (N>1: per exchange chunk); inputs already resident in HBM
                gms[s_].load(dev_batch["latents"], dev_batch["ctx"], dev_batch["pooled"], dev_batch["time_ids"],
                             t_embed[i * A + a], sig[i * A + a], None, 1.0 / A)
                gms[s_].replay(last=core.dp_last)
            else:
                core.step_no_autograd(grad_scale=1.0 / A, latents=dev_batch["latents"], ctx=dev_batch["ctx"],
                                      pooled=dev_batch["pooled"], time_ids=dev_batch["time_ids"], t_embed=t_embed[i * A + a],
                                      sig_or_t=sig[i * A + a], weight=None, loss_scale=1.0)
        core.dp_last = False
        trainer.optimizer_step()  # N>1: wait for the gradient exchange (or one NCCL all-reduce), then clip + AdamW
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    if use_graph:
        launches = sum(gms[s_].launches_per_replay for s_ in timed_seq) + K * og.launches_per_replay
    ms_dev = e0.elapsed_time(e1) / K

    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])

    # ---- (3) data-parallel proof (untimed): every rank must hold bit-identical parameters after the timed steps, and
    # bit-identical REDUCED gradients after one more exchanged micro-step (DDP's contract, src/core/distributed.py:142-163)
    dp_check = None
    if world > 1 and not args.no_grad_exchange:
        def checksum(buf):
            n2 = buf.numel() // 2 * 2
            a = buf[:n2].view(torch.int32).sum(dtype=torch.int64)
            b = buf.view(torch.int16).sum(dtype=torch.int64)  # a second, independent linear functional of the bits
            return torch.stack([a, b])

        sums = [checksum(unet.store.flat)]
        dev_batch = dev_batches[shapes[0]]
        core.dp_last = True
        if use_graph:
            gm.load(dev_batch["latents"], dev_batch["ctx"], dev_batch["pooled"], dev_batch["time_ids"], t_embed[0], sig[0],
                    None, 1.0)
            gm.replay(last=True)
        else:
            core.step_no_autograd(grad_scale=1.0, latents=dev_batch["latents"], ctx=dev_batch["ctx"],
                                  pooled=dev_batch["pooled"], time_ids=dev_batch["time_ids"], t_embed=t_embed[0],
                                  sig_or_t=sig[0], weight=None, loss_scale=1.0)
        core.dp_last = False
        if core.dp is not None:
            if not core.dp.issued:
                core.dp.exchange_all()
            core.dp.finish()
        else:
            allreduce_gradients(unet)
        torch.cuda.synchronize()
        sums.append(checksum(unet.store.grad))
        gnorm = unet.store.grad.float().norm().reshape(1)
        mine = torch.cat([torch.cat(sums), gnorm.to(torch.float64).view(torch.int64)])
        allsums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        unet.store.grad.zero_()
        same_p = all(torch.equal(a[:2], allsums[0][:2]) for a in allsums)
        same_g = all(torch.equal(a[2:4], allsums[0][2:4]) for a in allsums)
        dp_check = {"dp_params_identical": bool(same_p), "dp_reduced_grads_identical": bool(same_g),
                    "reduced_grad_l2": float(gnorm), "ranks_compared": world,
                    "how": "two 64-bit sums (over 32-bit and 16-bit words) of the flat bf16 parameter buffer after the timed steps, and of the flat "
                           "gradient buffer after one more exchanged micro-step, all-gathered and compared on every rank"}
        if not (same_p and same_g) or not float(gnorm) > 0.0:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "invalid": "data-parallel ranks diverged", **dp_check}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    if rank == 0 and core.dp is not None and core.dp.plan is not None:
        print(core.dp.plan.describe(), file=sys.stderr, flush=True)
    if rank == 0:
        step_flops = sum(train_step_flops(None, s_[0], s_[1]) * B for s_ in timed_seq) / K
        value = world * B * A / (ms_dev * 1e-3)
        e2e_v = world * B * A / (ms_e2e * 1e-3)
        roof = _time_gemm_roofline(ops, peaks)
        roof["step_tflops_per_gpu"] = round(step_flops / (ms_dev * 1e-3) / 1e12, 1)
        roof["step_frac_of_sustained_peak"] = round(step_flops / (ms_dev * 1e-3) / 1e12 / peaks["sustained"], 4)
        extra = _extra_rooflines(ops, peaks) if world == 1 else None
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            _, cb, _ = _cpu_baseline(timed_steps=2)
        out = {"metric": METRIC, "value": round(value, 4), "unit": "images/s", "n_gpus": world, "steps": K,
               "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev, 2), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": (WORKLOAD_FMT.format(method=args.method, W=8 * W, H=8 * H, h=H, w=W,
                                                           accum=(f"grad-accum {A}, " if A > 1 else ""), opt=args.optimizer)
                                       if not mixed else
                                       f"SDXL-base UNet, {args.method}, bs=4/GPU, mixed-AR buckets 768^2 / 1024^2 / 1280x960 (one "
                                       f"bucket per global micro-step, data.BucketBatchSampler), grad-accum {A}, bf16, full "
                                       f"fwd+bwd+loss+clip+{args.optimizer} (configs[4])"),
                          **({"bucket_counts": {f"{8 * s_[1]}x{8 * s_[0]}": timed_seq.count(s_) for s_ in shapes},
                              "images_1024eq_per_s": round(world * step_flops / train_step_flops(None, 128, 128)
                                                           / (ms_dev * 1e-3), 3),
                              "uncaptured_micro_steps": trainer.uncaptured_micro_steps} if mixed else {}),
                          "cuda_graph": use_graph,
                          "global_batch": B * world * A, "parallelism": f"dp{world}",
                          "grad_exchange": ("none (1 GPU)" if world == 1 else "DISABLED (diagnostic)" if args.no_grad_exchange else
                                            f"peer-memory exchange ({core.dp.mode}: reduce-scatter + all-gather over IPC-mapped "
                                            f"buffers) overlapped with backward, {core.dp.plan.n_chunks} chunks"
                                            if core.dp is not None and core.dp.plan
                                            else "one NCCL all-reduce of the flat gradient buffer after backward"),
                          "l2": "working set >> L2: 5.1 GB of weights + ~40 GB activations streamed every step",
                          "last_loss": last_loss,
                          "last_loss_note": "random-init weights under the zero-terminal-SNR schedule (sigma up to 2e4): the "
                                            "reference clamps the loss at 1000 (ddpm_trainer.py:380-384)"},
               "e2e": {"value": round(e2e_v, 4), "unit": "images/s", "h2d_bytes_per_step": int(h2d) * A,
                       "d2h_bytes_per_step": (4 + 6 * 8) * A, "ms_per_step": round(ms_e2e, 2),
                       "api": "B200DDPMTrainer._execute_training_step(batch) with pinned host tensors"
                              + (" (cuda_graph=True)" if use_graph else "")},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_extra": extra,
               **(dp_check or {}),
               **({"invalid": "diagnostic run without gradient exchange (independent replicas)"}
                  if args.no_grad_exchange and world > 1 else {}),
               "cpu_baseline": cb, "torch_eager_gpu": eager}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_resnet320(args):
    """BASELINE.json configs[0] / BASELINE.md B-CPU-1: single ResnetBlock2D (320 -> 320 channels, 64x64 latent, B=4),
    forward + backward on the kernels (CUDA-graph replay, CUDA events) next to the oracle's ResnetBlock2D in PyTorch eager
    fp32 on the host cores, with the loss match on identical bf16-rounded weights / inputs.  Prints one JSON line."""
    from oracle.unet_sdxl import ResnetBlock2D
    from sdxl_training_improvements_b200.modules import ModuleRunner, resnet_block_flops
    torch.cuda.set_device(0)
    peaks = _peaks()
    B, H, W, Cc, temb = 4, 64, 64, 320, 1280
    torch.manual_seed(0)
    ref = ResnetBlock2D(Cc, Cc, temb)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.copy_(p_.to(torch.bfloat16).float())
    x = torch.randn(B, Cc, H, W).to(torch.bfloat16).float()
    emb = torch.randn(B, temb).to(torch.bfloat16).float()
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)

    def cpu_step():
        xr, er = x.clone().requires_grad_(True), emb.clone().requires_grad_(True)
        t0 = time.time()
        loss = ref(xr, er).square().mean()
        loss.backward()
        return time.time() - t0, float(loss)

    cpu_step()
    cpu_t = sorted(cpu_step()[0] for _ in range(max(5, args.steps)))
    cpu_ms = cpu_t[len(cpu_t) // 2] * 1e3
    loss_ref = cpu_step()[1]

    run = ModuleRunner(Cc)
    pfx = "down_blocks.0.resnets.0"
    run.load_module_state(pfx, ref.state_dict())
    xg, eg = x.cuda(), emb.cuda()
    out = run.resnet_forward(pfx, xg, eg)
    loss_k = float(out.float().square().mean())
    dout = ((2.0 / out.numel()) * out.float()).contiguous()
    run.resnet_backward(dout)
    # timed region = the module's kernels only, token-major operands resident (what the UNet sees inside a step)
    xa = run._to_tokens(xg)
    ea = eg.to(torch.bfloat16).contiguous()
    da = run._to_tokens(dout)
    from sdxl_training_improvements_b200.unet import Act

    def fwd_bwd():
        run.eng.tape = []
        xi, ei = Act(xa), Act(ea)
        o = run.eng.resnet(xi, run.eng.silu(ei), B, H, W, Cc, Cc, pfx)
        o.g = da
        run._run_tape_backward()

    def fwd_only():
        run.eng.tape = []
        run.eng.resnet(Act(xa), run.eng.silu(Act(ea)), B, H, W, Cc, Cc, pfx)
        run.eng.tape = []

    iters = max(10, args.steps)
    us_fb = _graph_time_us(fwd_bwd, iters=iters)
    us_f = _graph_time_us(fwd_only, iters=iters)
    f_fwd = resnet_block_flops(B, H, W, Cc, Cc, temb)
    out_line = {
        "metric": "resnet_block_320_64x64_fwd_bwd_us", "value": round(us_fb, 1), "unit": "us", "n_gpus": 1,
        "steps": iters, "warmup": 2, "higher_is_better": False, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "single ResnetBlock2D 320->320, 64x64 latent, B=4, fwd+bwd (BASELINE configs[0])"},
        "fwd_us": round(us_f, 1), "fwd_gflop": round(f_fwd / 1e9, 2), "fwd_bwd_gflop": round(3 * f_fwd / 1e9, 2),
        "roofline": {"bound": "tensor", "achieved": round(3 * f_fwd / us_fb / 1e6, 1), "peak": peaks["burst"],
                     "unit": "TFLOP/s", "frac": round(3 * f_fwd / us_fb / 1e6 / peaks["burst"], 4), "traffic": None,
                     "note": "45.3 GFLOP per sample fwd+bwd (SURVEY 8d conv roofline figure) x B=4; the block is 2 convs at "
                             "Cin=320 (5 k-blocks per tap) + 2 GroupNorm+SiLU passes + time-embedding row"},
        "loss_kernel": loss_k, "loss_cpu_oracle": loss_ref, "loss_abs_diff": abs(loss_k - loss_ref),
        "cpu_baseline": {"value": round(cpu_ms, 1), "unit": "ms", "cores": ncores, "kind": "port",
                         "sample": "oracle ResnetBlock2D 320->320 @64x64 B=4, PyTorch eager fp32 on the host cores, fwd+bwd, "
                                   f"median of {len(cpu_t)}"},
        "speedup_vs_cpu": round(cpu_ms * 1e3 / us_fb, 1)}
    print(json.dumps(out_line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--method", default="ddpm", choices=["ddpm", "flow_matching"])
    ap.add_argument("--latent-h", type=int, default=128, help="latent height (image / 8); default 128 = 1024 px")
    ap.add_argument("--latent-w", type=int, default=128)
    ap.add_argument("--accum", type=int, default=1, help="gradient-accumulation micro-steps per optimizer step")
    ap.add_argument("--mixed-buckets", action="store_true",
                    help="BASELINE configs[4]: aspect buckets {768^2, 1024^2, 1280x960}, one bucket per global micro-step "
                         "(use with --method flow_matching --accum 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU oracle timing")
    ap.add_argument("--optimizer", default="adamw_bf16", choices=["adamw_bf16", "adamw_fp32"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of CUDA graphs")
    ap.add_argument("--no-grad-exchange", action="store_true",
                    help="DIAGNOSTIC (N>1): skip the gradient exchange altogether — independent replicas, the max-over-ranks "
                         "time then shows what the slowest GPU of the box costs; the JSON line is marked invalid")
    ap.add_argument("--config", default="step", choices=["step", "resnet320"],
                    help="step: the SDXL training step (BASELINE configs[1..4]); resnet320: BASELINE configs[0], one "
                         "ResnetBlock2D 320ch @64x64 kernel-vs-CPU-eager with loss match")
    args = ap.parse_args()
    if args.config == "resnet320":
        run_resnet320(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
