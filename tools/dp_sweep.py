"""
Data-parallel schedule sweep (run under torch.distributed.run, one rank per GPU): ONE process group, one communicator, and the
schedule knobs of the fused step changed at run time between short timed runs, so that a whole grid costs seconds of GPU time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dp_sweep.py --out gpurun_out/dp_sweep_n8.txt

Knobs (environment, read by the library on every call): SLICED_DP_CHUNKS, SLICED_DP_CHUNKS_ALL, SLICED_DP_ORDER, SLICED_DP_LAYER_SGD,
SLICED_GEMM_RESERVE_SMS, SLICED_GEMM_MAX_SPLITS.  Every rank sets the same values in the same order (collectives stay matched).
Timing: CUDA events around K steps after a barrier, max over ranks; configurations are interleaved round-robin over PASSES passes
(clock / thermal drift).  Also times the bare NCCL all-reduce of one 64 MB weight gradient and of the whole 134 MB bucket.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench as BN

KNOBS = ("SLICED_DP_NOCOMM", "SLICED_DP_CHUNKS", "SLICED_DP_CHUNKS_ALL", "SLICED_DP_ORDER", "SLICED_DP_LAYER_SGD", "SLICED_GEMM_RESERVE_SMS", "SLICED_GEMM_MAX_SPLITS",
         "SLICED_GEMM_SCHED")
CONFIGS = [
    ("static schedule, chunks 4 (r2 default)", dict(SLICED_GEMM_SCHED="0", SLICED_DP_CHUNKS="4")),
    ("static schedule, chunks 1", dict(SLICED_GEMM_SCHED="0", SLICED_DP_CHUNKS="1")),
    ("launch control, chunks 1", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="1")),
    ("launch control, chunks 2", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="2")),
    ("launch control, chunks 4", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="4")),
    ("launch control, chunks 4, every layer", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="4", SLICED_DP_CHUNKS_ALL="1")),
    ("launch control, chunks 1, dX-first order", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="1", SLICED_DP_ORDER="1")),
    ("launch control, chunks 1, flat SGD", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="1", SLICED_DP_LAYER_SGD="0")),
    ("launch control, chunks 1, no split-K", dict(SLICED_GEMM_SCHED="1", SLICED_DP_CHUNKS="1", SLICED_GEMM_MAX_SPLITS="1")),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--scaling", default="strong")
    args = ap.parse_args()
    import sliced_b200 as S
    from sliced_b200 import capi, dp
    from sliced_b200.host import CUDA, Mlp
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    dev = CUDA(local, cached=True, stream=stream.cuda_stream)
    lib, ctx = capi.load(), dev.ctx_handle
    dp.init_comm(lib, ctx, dist, rank, world)
    batch = BN.GLOBAL_BATCH // world if args.scaling == "strong" else BN.GLOBAL_BATCH
    gb = batch * world
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    x = torch.empty(batch, BN.DIMS[0], device="cuda").uniform_(0, 1, generator=g)
    lab = torch.randint(0, BN.DIMS[-1], (batch,), device="cuda", generator=g, dtype=torch.int32)
    y = torch.zeros(batch, BN.DIMS[-1], device="cuda"); y[torch.arange(batch, device="cuda"), lab.long()] = 1.0
    mlp = Mlp(dev, BN.DIMS, 0)
    mlp.set_fused(True)
    W, B = BN.make_params()
    for l in range(3):
        mlp.weights(l).write(W[l]); mlp.bias(l).write(B[l])
    bx, by = dev.wrap(x.data_ptr(), x.numel()).no_grad(), dev.wrap(y.data_ptr(), y.numel()).no_grad()
    bl = dev.wrap(lab.data_ptr(), batch, np.int32)

    def set_knobs(kv):
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update(kv)

    def timed(fn, n):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
        tmin = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        return float(t.item()), float(tmin.item())

    step = lambda: mlp.step(bx, by, bl, batch, BN.LR, grad_rows=gb, want_metrics=False)
    res = {name: [] for name, _ in CONFIGS}
    for p in range(args.passes):
        for name, kv in CONFIGS:
            set_knobs(kv)
            for _ in range(3):
                step()
            res[name].append(timed(step, args.steps)[0])
    set_knobs({})
    nccl_env = {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}
    # the step without its exchange (wrong numbers, right cost): the per-GPU compute share, and how far apart the ranks run
    os.environ["SLICED_DP_NOCOMM"] = "1"
    for _ in range(3):
        step()
    t_nc = timed(step, args.steps)
    os.environ.pop("SLICED_DP_NOCOMM")
    # bare exchange cost
    bucket = mlp.grad_bucket()
    nW = BN.DIMS[0] * BN.DIMS[1]
    ar_w = lambda: capi.check(ctx, lib.sl_allreduce_sum(ctx, S.F32, C.c_void_p(bucket.ptr), nW))
    ar_all = lambda: capi.check(ctx, lib.sl_allreduce_sum(ctx, S.F32, C.c_void_p(bucket.ptr), mlp.n_params))
    for f in (ar_w, ar_all):
        for _ in range(3):
            f()
    t_w, t_all = timed(ar_w, 20), timed(ar_all, 20)
    lines = [f"# NCCL environment: {nccl_env}",
             f"# data-parallel schedule sweep, N={world}, per-GPU batch {batch} ({args.scaling}), {args.steps} steps x {args.passes} passes, ms/step = max over ranks",
             f"# bare NCCL all-reduce (alone on the GPU): one weight gradient {nW * 4 / 1e6:.0f} MB {t_w[0]:.3f} ms; whole bucket {mlp.n_params * 4 / 1e6:.0f} MB {t_all[0]:.3f} ms",
             f"# step without any exchange (per-GPU compute only): slowest rank {t_nc[0]:.3f} ms, fastest rank {t_nc[1]:.3f} ms",
             f"# {'configuration':40s} " + " ".join(f"pass{p}" for p in range(args.passes)) + "    best"]
    for name, _ in CONFIGS:
        v = res[name]
        lines.append(f"{name:42s} " + " ".join(f"{t:6.3f}" for t in v) + f"  {min(v):6.3f}")
    if rank == 0:
        txt = "\n".join(lines)
        print(txt, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(txt + "\n")
    del mlp
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
