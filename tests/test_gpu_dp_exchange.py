"""Data-parallel gradient exchange over peer memory on real GPUs (needs >= 2; skipped otherwise).

  * transport: random bf16 buffers, a multi-chunk plan, chunks exchanged one by one -> every rank holds
    bf16(exact sum) bit for bit (fp32 accumulation of <= 16 bf16 values of similar magnitude is exact), memory outside
    the exchanged chunks untouched; repeated for several sequence numbers (flag reuse);
  * trainer: tiny UNet, the same seeds, three optimizer steps (one with gradient accumulation) through
    `_execute_training_step` with the overlapped peer exchange and with ONE NCCL all-reduce after backward
    (B2_DP_EXCHANGE=nccl), each both eager and with CUDA graphs (one graph per chunk).  Parameters after the steps must
    be bit-identical across ranks in every mode; peer vs NCCL must agree to rel-L2 <= 2e-2 on the parameter UPDATE (the
    two-term bf16 sums are identical, but the fp32 atomics that stage bias / norm gradients make any two runs differ in
    the last bit, and AdamW's first steps are sign-like), losses to 1e-3.
Reference behaviour being replaced: DistributedDataParallel's gradient all-reduce, src/core/distributed.py:142-163.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    return dist


def _transport_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    dist = _init(rank, world, port)
    try:
        from sdxl_training_improvements_b200 import dp as D
        total = 8 * 300_000 + 8 * 13
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        grad = torch.zeros(total, device="cuda", dtype=torch.bfloat16)
        x = D.PeerGradExchange(grad, self_test=True, mode="auto", autotune=False)  # all three transports
        # three chunks, one of them in two pieces, one tiny, plus a region nobody exchanges
        chunks = [[(0, 8 * 100_000)], [(8 * 100_000, 8 * 7), (8 * 150_000, 8 * 100_000)], [(8 * 250_000, 8 * 40_000)]]
        base = 0
        for k, rg in enumerate(chunks):
            base = x._set_chunk(k, rg, base)
        inside = torch.zeros(total, dtype=torch.bool, device="cuda")
        for rg in chunks:
            for o, n in rg:
                inside[o:o + n] = True
        ok = True
        msgs = []
        for it in range(6):
            x.mode = list(x.handles)[it % len(x.handles)]
            mine = (torch.randn(total, device="cuda", generator=g) * (1 + rank)).to(torch.bfloat16)
            grad.copy_(mine)
            everyone = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            want = sum(e.double() for e in everyone).to(torch.bfloat16)
            torch.cuda.synchronize()
            dist.barrier()
            for k in range(len(chunks)):
                x.exchange_chunk(k)
            x.finish()
            torch.cuda.synchronize()
            bad_in = int((grad[inside] != want[inside]).sum())
            bad_out = int((grad[~inside] != mine[~inside]).sum())
            if bad_in or bad_out:
                ok = False
                msgs.append(f"iter {it} ({x.mode}): {bad_in} wrong sums, {bad_out} elements outside the chunks changed")
            dist.barrier()
        # whole buffer, every transport
        for m in x.handles:
            x.mode = m
            mine = torch.randn(total, device="cuda", generator=g).to(torch.bfloat16)
            grad.copy_(mine)
            ref = mine.clone()
            dist.all_reduce(ref)
            torch.cuda.synchronize()
            dist.barrier()
            x.exchange_all()
            x.finish()
            torch.cuda.synchronize()
            if world == 2 and not torch.equal(grad, ref):
                ok = False
                msgs.append(f"exchange_all ({m}) differs from the NCCL all-reduce at world size 2")
            dist.barrier()
        x.autotune()
        x.close()
        q.put((rank, ok, msgs))
    finally:
        dist.destroy_process_group()


def _trainer_worker(rank, world, port, q, mode):
    sys.path.insert(0, ROOT)
    if mode.endswith("nccl"):
        os.environ["B2_DP_EXCHANGE"] = "nccl"
    dist = _init(rank, world, port)
    try:
        from types import SimpleNamespace
        from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
        from sdxl_training_improvements_b200.trainer import B200AdamW, B200DDPMTrainer
        from sdxl_training_improvements_b200.unet import B200UNet
        cfg = tiny_config()
        ref = seeded_init_(OracleUNet(cfg), 0)
        net = B200UNet(cfg, device=f"cuda:{rank}")
        net.load_state_dict(ref.state_dict())
        opt = B200AdamW(net, lr=1e-3, weight_decay=0.0)
        conf = SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0,
                                                     use_ztsnr=True, min_snr_gamma=None),
                               training=SimpleNamespace(method="ddpm", prediction_type="v_prediction",
                                                        gradient_accumulation_steps=2, clip_grad_norm=1.0))
        tr = B200DDPMTrainer(net, opt, None, f"cuda:{rank}", config=conf, seed=10 + rank, cuda_graph=mode.startswith("graph"))
        used_peer = tr.core.dp is not None
        B, H, W = 2, 16, 16
        gen = torch.Generator().manual_seed(500 + rank)
        losses = []
        for step in range(3):
            A = 2 if step == 1 else 1
            for a in range(A):
                batch = {"vae_latents": torch.randn(B, 4, H, W, generator=gen),
                         "prompt_embeds": torch.randn(B, 77, cfg["cross_attention_dim"], generator=gen),
                         "pooled_prompt_embeds": torch.randn(B, 96, generator=gen),
                         "time_ids": torch.tensor([[128., 128., 0., 0., 128., 128.]]).repeat(B, 1)[:, None],
                         "metadata": [{} for _ in range(B)]}
                torch.manual_seed(1000 * step + 10 * a + rank)  # timestep draw (host RNG) identical across modes
                loss, _ = tr._execute_training_step(batch, accumulate=A > 1, is_last_accumulation_step=a == A - 1)
                losses.append(float(loss))
        torch.cuda.synchronize()
        flat = net.store.flat.clone()
        init = B200UNet(cfg, device=f"cuda:{rank}")
        init.load_state_dict(ref.state_dict())
        delta = (flat.float() - init.store.flat.float()).cpu()
        everyone = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(everyone, flat)
        same = all(torch.equal(everyone[0], e) for e in everyone)
        n_chunks = tr.core.dp.plan.n_chunks if used_peer and tr.core.dp.plan is not None else 0
        q.put((rank, mode, used_peer, same, losses, delta if rank == 0 else None, n_chunks))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _equivalence_worker(rank, world, port, q):
    """T6: N ranks x batch b == 1 rank x batch N*b (DDP's contract, src/core/distributed.py:142-163: gradients are the
    MEAN over ranks of per-rank mean-loss gradients).  Global batch of 4 samples with explicit noise / timesteps; rank r
    takes samples [2r, 2r+2) through the plugin + the overlapped peer exchange; every rank also runs all 4 samples on
    its own GPU without any exchange.  reduced_gradient / world must equal the single-GPU gradient."""
    sys.path.insert(0, ROOT)
    dist = _init(rank, world, port)
    try:
        from types import SimpleNamespace
        from oracle.unet_sdxl import OracleUNet, seeded_init_, tiny_config
        from sdxl_training_improvements_b200.trainer import B200AdamW, B200DDPMTrainer
        from sdxl_training_improvements_b200.unet import B200UNet
        cfg = tiny_config()
        sd = seeded_init_(OracleUNet(cfg), 0).state_dict()
        conf = SimpleNamespace(model=SimpleNamespace(num_timesteps=1000, sigma_min=0.002, sigma_max=20000.0,
                                                     use_ztsnr=True, min_snr_gamma=5.0),
                               training=SimpleNamespace(method="ddpm", prediction_type="v_prediction",
                                                        gradient_accumulation_steps=1, clip_grad_norm=1.0))
        nets, trs = [], []
        for _ in range(2):
            net = B200UNet(cfg, device=f"cuda:{rank}")
            net.load_state_dict(sd)
            nets.append(net)
            trs.append(B200DDPMTrainer(net, B200AdamW(net, lr=1e-3), None, f"cuda:{rank}", config=conf, seed=3))
        tr_dp, tr_one = trs
        tr_one.core.dp, tr_one.world_size = None, 1          # the single-GPU twin: no exchange
        used_peer = tr_dp.core.dp is not None
        Bg, H, W = 4, 16, 16
        g = torch.Generator().manual_seed(77)                 # the same global batch on every rank
        full = {"vae_latents": torch.randn(Bg, 4, H, W, generator=g),
                "prompt_embeds": torch.randn(Bg, 77, cfg["cross_attention_dim"], generator=g),
                "pooled_prompt_embeds": torch.randn(Bg, 96, generator=g),
                "time_ids": torch.tensor([[128., 128., 0., 0., 128., 128.]]).repeat(Bg, 1)[:, None],
                "metadata": [{} for _ in range(Bg)]}
        noise = torch.randn(Bg, 4, H, W, generator=g)
        ts = torch.tensor([650, 720, 800, 880])
        b = Bg // world
        sl = slice(rank * b, (rank + 1) * b)
        mine = {k: (v[sl] if torch.is_tensor(v) else v[sl]) for k, v in full.items()}
        rels, norm_checks = [], []
        for rep in range(2):   # pass 0: no chunk plan yet -> whole-buffer exchange; pass 1: chunks leave during backward
            for n_ in nets:
                n_.zero_grad()
            tr_dp.core.dp_last = True
            out = tr_dp.training_step(mine, noise=noise[sl], timesteps=ts[sl])
            out["loss"].backward()
            tr_dp.core.dp_last = False
            if used_peer:
                x = tr_dp.core.dp
                if not x.issued:
                    x.exchange_all()
                if x.provides_norm:  # the exchange's reduce kernels also deliver sum((reduced gradient)^2)
                    gn = torch.full((1,), -1.0, device="cuda", dtype=torch.float64)
                    x.finish(gnorm_sq_out=gn)
                else:
                    gn = None
                    x.finish()
            else:
                gn = None
                dist.all_reduce(nets[0].store.grad)
            if gn is not None:
                from sdxl_training_improvements_b200 import ops
                chk = torch.zeros(1, device="cuda", dtype=torch.float64)
                ops.sumsq(nets[0].store.grad, chk)
                torch.cuda.synchronize()
                gns = [torch.empty_like(gn) for _ in range(world)]
                dist.all_gather(gns, gn)
                norm_checks.append((abs(float(gn) - float(chk)) / max(float(chk), 1e-30), all(torch.equal(gns[0], t) for t in gns)))
            out1 = tr_one.training_step(full, noise=noise, timesteps=ts)
            out1["loss"].backward()
            torch.cuda.synchronize()
            red = nets[0].store.grad.float() / world
            one = nets[1].store.grad.float()
            rels.append(float((red - one).norm() / one.norm()))
            lsum = torch.tensor([float(out["loss"])], device="cuda")
            dist.all_reduce(lsum)
            loss_dp, loss_one = float(lsum) / world, float(out1["loss"])
        flat = nets[0].store.grad.clone()
        everyone = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(everyone, flat)
        same = all(torch.equal(everyone[0], e) for e in everyone)
        q.put((rank, used_peer, rels, loss_dp, loss_one, same, norm_checks))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _spawn(target, world, extra=()):
    import torch.multiprocessing as mp
    port = 32500 + (os.getpid() % 2000) + len(extra) * 7 + (hash(extra) % 50)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + tuple(extra)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda r: r[0])


def _world():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    return min(n, 4)


@pytest.mark.timeout(600)
def test_peer_exchange_transport_bit_exact():
    world = _world()
    for rank, ok, msgs in _spawn(_transport_worker, world):
        assert ok, f"rank {rank}: {msgs}"


@pytest.mark.timeout(900)
def test_trainer_overlapped_exchange_matches_nccl_allreduce():
    world = 2
    _world()
    out = {}
    for mode in ("peer", "nccl", "graph", "graph_nccl"):
        res = _spawn(_trainer_worker, world, (mode,))
        for rank, m, used_peer, same, losses, delta, n_chunks in res:
            assert same, f"{mode}: parameters differ across ranks after 3 steps"
            assert used_peer == (not mode.endswith("nccl")), f"{mode}: peer exchange in use = {used_peer}"
            if not mode.endswith("nccl"):
                assert n_chunks >= 2, f"{mode}: the plan has {n_chunks} chunk(s) — no overlap possible"
        out[mode] = res[0]
    for mode, base in (("peer", "nccl"), ("graph", "graph_nccl")):
        d, r = out[mode][5], out[base][5]
        rel = float((d - r).norm() / r.norm())
        assert rel <= 2e-2, f"{mode}: parameter update differs from the NCCL all-reduce path, rel-L2 {rel:.3e}"
        assert all(abs(a - b) <= 1e-3 * max(1.0, abs(b)) for a, b in zip(out[mode][4], out[base][4])), \
            (out[mode][4], out[base][4])


@pytest.mark.timeout(600)
def test_two_ranks_times_b_equals_one_rank_times_2b():
    """T6 (VERDICT r1 missing #4).  Tolerances (stated): gradient rel-L2 <= 1e-2 (two bf16 partial sums vs one bf16 sum of
    four samples: different rounding points, same fp32 mathematics), loss |d| <= 1e-3 * max(1, loss)."""
    _world()
    for rank, used_peer, rels, loss_dp, loss_one, same, norm_checks in _spawn(_equivalence_worker, 2):
        assert used_peer, "the peer-memory exchange was not in use"
        for rel, identical in norm_checks:  # norm folded into the exchange == b2_sumsq over the reduced buffer, same bits on all ranks
            assert rel <= 1e-5 and identical, f"rank {rank}: exchange-provided grad norm^2 off by {rel:.2e}, identical={identical}"
        assert same, "reduced gradients differ across ranks"
        assert all(r <= 1e-2 for r in rels), f"rank {rank}: reduced gradient / world vs single-GPU gradient rel-L2 {rels}"
        assert abs(loss_dp - loss_one) <= 1e-3 * max(1.0, abs(loss_one)), (loss_dp, loss_one)


if __name__ == "__main__":
    w = min(torch.cuda.device_count(), 4)
    print("transport:", _spawn(_transport_worker, w))
    res = {}
    for mode in ("peer", "nccl", "graph", "graph_nccl"):
        r = _spawn(_trainer_worker, 2, (mode,))
        res[mode] = r[0]
        print(mode, [(x[0], x[2], x[3], x[4], x[6]) for x in r], flush=True)
    for mode, base in (("peer", "nccl"), ("graph", "graph_nccl")):
        d, r = res[mode][5], res[base][5]
        print(mode, "vs", base, "update rel-L2", float((d - r).norm() / r.norm()), flush=True)
