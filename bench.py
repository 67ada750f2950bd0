#!/usr/bin/env python
"""
bench.py — headline benchmark of sliced_b200 (contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[4], the config its metric "MLP train-step samples/s @1/2/4/8 GPU" is quoted on; it fits
one GPU): examples/nn.rs MLP 4096-4096-4096-10, softmax + categorical cross-entropy, SGD lr 0.1, f32, GLOBAL batch 65536,
data-parallel along the batch with one NCCL sum-all-reduce of the 134 MB gradient bucket per step (strong scaling).
One "step" = zero_grad -> 3 x (gemm, add_row_mut, relu|softmax) -> accuracy + loss -> backward_with (tape) -> all-reduce ->
SGD, exactly the op sequence of examples/nn.rs:184-237, run through the C++ host layer over the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak] [--gemm-mode 3xtf32|tf32|3xf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

value   whole-job samples/s with the batch already resident in HBM (device-timed with CUDA events, max over ranks)
e2e     the same metric through the host-layer API with HOST buffers: every step copies its inputs from pinned host memory
        and copies the step's loss/accuracy back to (pinned) host memory
roofline      the dominant kernel (tcgen05 gemm): per-launch CUDA events recorded live during the timed steps
cpu_baseline  the reference's CPU path (oracle replay + OpenBLAS sgemm, what custos' `blas` feature links) on the host cores,
              on a bounded sample of the same workload — a reported baseline, not the target
--impl reference   times that CPU path alone (rank 0 only), same metric / config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

DIMS = [4096, 4096, 4096, 10]
GLOBAL_BATCH = 65536
LR = 0.1
METRIC = "mlp_train_step_samples_per_s"
UNIT = "samples/s"
WORKLOAD = "examples/nn.rs MLP 4096-4096-4096-10 softmax+cce SGD(lr 0.1) f32, global batch 65536, data-parallel over the batch"


def step_flops(batch: int) -> float:
    """algorithmic gemm flops of one training step: fwd 2B(d0 d1 + d1 d2 + d2 d3), bwd dW for all layers + dA for layers > 0"""
    f = 0.0
    for l in range(len(DIMS) - 1):
        f += 2.0 * batch * DIMS[l] * DIMS[l + 1]          # forward
        f += 2.0 * batch * DIMS[l] * DIMS[l + 1]          # dW
        if l > 0:
            f += 2.0 * batch * DIMS[l] * DIMS[l + 1]      # dA (x is no_grad: layer 0 has none)
    return f


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
    of this very command (profiles/r1_ncu_traffic.json, written from the .ncu-rep by tools/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    try:
        d = json.load(open(p))
        return dict(bytes_per_launch=d["dram_bytes_per_launch"], unit="B", kernel=d["kernel"], algorithmic_bytes_per_launch=d.get("algorithmic_bytes_per_launch"),
                    launches_averaged=d.get("launches_captured"), source=d.get("source"))
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index):
    """Run this rank's host threads on the NUMA node its GPU hangs off (sysfs), so that pinned host buffers are allocated there and
    host->device copies do not cross sockets (8 ranks uploading 1 GB per step otherwise share one inter-socket link).  Returns the
    node, or None when the topology cannot be read (then nothing is changed).  SLICED_BENCH_NUMA=0 switches it off."""
    if os.environ.get("SLICED_BENCH_NUMA", "1") == "0":
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # samples that arrived inside the timed region; a region shorter than a few sampling periods (multi-GPU strong scaling:
        # ~70 ms) falls back to every sample since the sampler started, i.e. warm-up steps (same kernels, same load) + timed region
        inside = [ln for (t, ln) in self.lines if self.t_begin and self.t_end and self.t_begin <= t <= self.t_end + 0.05]
        window = "timed region"
        if len(inside) < 3:
            inside = [ln for (_, ln) in self.lines]
            window = "warm-up + timed region (timed region shorter than 3 sampling periods)"
        sm, mx, reasons, pw = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        busy = [c for c, w in zip(sm, pw) if w > 0.5 * max(pw)] or sm   # under load: drop idle samples from before the first kernel
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), power_w_max=max(pw), samples=len(busy),
                    window=window)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def make_host_problem(batch: int, seed: int = 7):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (batch, DIMS[0])).astype(np.float32)      # shipped data is pixel/255 (nn.rs:170-173)
    labels = rng.integers(0, DIMS[-1], batch).astype(np.int32)
    y = np.zeros((batch, DIMS[-1]), np.float32)
    y[np.arange(batch), labels] = 1.0
    return x, y, labels


def make_params(seed: int = 11):
    rng = np.random.default_rng(seed)
    # nn.rs:26 draws U(-0.1, 0.1); at width 4096 that saturates the softmax in the first forward pass and the reference's own
    # cce_grad (targets / preds, nn.rs:150) then divides 0 by 0.  The bound is therefore capped at the Xavier limit so that the
    # synthetic run trains (finite loss); shapes, op sequence and cost are unchanged.
    W = [rng.uniform(-1, 1, DIMS[i] * DIMS[i + 1]).astype(np.float32) * np.float32(min(0.1, (6.0 / (DIMS[i] + DIMS[i + 1])) ** 0.5))
         for i in range(len(DIMS) - 1)]
    B = [np.zeros(DIMS[i + 1], np.float32) for i in range(len(DIMS) - 1)]                                   # nn.rs:33
    return W, B


def cpu_reference_rate(steps: int, warmup: int, sample_batch: int):
    """The reference's own CPU implementation of the step: the oracle's op-by-op replay with OpenBLAS sgemm on all host cores."""
    import oracle as O
    cores = os.cpu_count() or 1
    blas = O.use_openblas(cores)
    x, y, labels = make_host_problem(sample_batch)
    W, B = make_params()
    xs, ys = x.ravel(), y.ravel()
    for _ in range(warmup):
        O.mlp_step(0, DIMS, xs, ys, labels, W, B, LR, grad_rows=GLOBAL_BATCH)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.mlp_step(0, DIMS, xs, ys, labels, W, B, LR, grad_rows=GLOBAL_BATCH)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    O.use_naive_gemm()
    return dict(value=sample_batch / dt, unit=UNIT, cores=cores if blas else 1, kind="port",
                sample=f"{steps} step(s) of the same MLP on a {sample_batch}-sample batch ({dt * 1e3:.0f} ms/step); "
                       f"gemm = OpenBLAS sgemm on {cores if blas else 1} thread(s), every other op = single-thread oracle loop "
                       f"(the reference's slice loops are single-threaded)"), dt


def cpu_sine_net_us(steps=300):
    """examples/sine_net.rs at its shipped sizes on the CPU port: us per training step.  gemm = OpenBLAS sgemm on ONE thread (measured
    here: 1.8 ms per step; 16 threads thrash on 1000 x 64 matrices: 47 ms; the naive loop: 6.9 ms), everything else the single-thread
    oracle loops like the reference's."""
    import oracle as O
    O.use_openblas(1)
    dims = [1, 64, 64, 1]
    xs = (np.arange(1000) / 1000.0).astype(np.float32)
    ys = np.sin(2.0 * xs * np.float32(np.pi)).astype(np.float32)
    rng = np.random.default_rng(0)
    W = [rng.uniform(-0.5, 0.5, dims[i] * dims[i + 1]).astype(np.float32) for i in range(3)]
    B = [np.zeros(dims[i + 1], np.float32) for i in range(3)]
    for _ in range(10):
        O.mlp_step(1, dims, xs, ys, None, W, B, 1e-4)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.mlp_step(1, dims, xs, ys, None, W, B, 1e-4)
    us = round(1e6 * (time.perf_counter() - t0) / steps, 1)
    O.use_naive_gemm()
    return us


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the FULL workload (global batch 65536, ~15-20 s of host time per step) on every host core: a bounded number of steps keeps the
    # run inside a few minutes; --cpu-sample N shrinks the batch (then config.sample_batch says so)
    steps, warmup = max(1, min(args.steps, 3 if args.cpu_sample >= GLOBAL_BATCH else 5)), min(args.warmup, 1)
    cb, dt = cpu_reference_rate(steps, warmup, args.cpu_sample)
    line = dict(metric=METRIC, value=cb["value"], unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, global_batch=GLOBAL_BATCH, sample_batch=args.cpu_sample, parallelism="cpu"),
                cpu_baseline=cb, e2e=dict(value=cb["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    rc_fail = False
    import torch
    import torch.distributed as dist

    import sliced_b200 as S
    from sliced_b200 import capi
    from sliced_b200.host import CUDA, Mlp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    # (multi-GPU runs only: at N=1 nothing competes for the host links, and the CPU-baseline leg keeps all host cores)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None   # before any pinned allocation: first touch on that node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stream = torch.cuda.current_stream()
    dev = CUDA(local_rank, cached=True, stream=stream.cuda_stream)
    dev.set_gemm_mode({"tf32": S.GEMM_TF32, "3xtf32": S.GEMM_3XTF32, "3xf16": S.GEMM_3XF16}[args.gemm_mode])
    lib = capi.load()
    ctx = dev.ctx_handle

    if world > 1:  # one NCCL communicator for the gradient exchange, id shipped through torch.distributed
        from sliced_b200 import dp
        dp.init_comm(lib, ctx, dist, rank, world)

    batch = GLOBAL_BATCH // world if args.scaling == "strong" else GLOBAL_BATCH
    if args.debug_per_gpu_batch:   # diagnosis only (e.g. the per-GPU share of an 8-GPU run on one GPU); the line says so in config
        batch = args.debug_per_gpu_batch
    global_batch = batch * world
    # synthetic data of the config's shape: this rank's shard (seeded per rank), pinned on the host for the e2e leg
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    x_dev = torch.empty(batch, DIMS[0], device="cuda").uniform_(0, 1, generator=g)
    labels_dev = torch.randint(0, DIMS[-1], (batch,), device="cuda", generator=g, dtype=torch.int32)
    y_dev = torch.zeros(batch, DIMS[-1], device="cuda")
    y_dev[torch.arange(batch, device="cuda"), labels_dev.long()] = 1.0
    x_pin = torch.empty(batch, DIMS[0], pin_memory=True).copy_(x_dev)
    y_pin = torch.empty(batch, DIMS[-1], pin_memory=True).copy_(y_dev)
    l_pin = torch.empty(batch, dtype=torch.int32, pin_memory=True).copy_(labels_dev)
    torch.cuda.synchronize()

    mlp = Mlp(dev, DIMS, 0)
    mlp.set_fused(not args.unfused)
    # data-parallel runs pipeline each layer's gradient join + SGD update into the next step's forward pass (bit-identical; see
    # Mlp::set_deferred); every timed region below ends with mlp.flush(), so all of its updates are inside it.  SLICED_DP_DEFERRED=0: off
    deferred = world > 1 and not args.unfused and os.environ.get("SLICED_DP_DEFERRED", "1") != "0"
    mlp.set_deferred(deferred)
    W, B = make_params()  # identical seeded init on every rank (replicas stay bit-identical: same summed gradients)
    for l in range(len(DIMS) - 1):
        mlp.weights(l).write(W[l])
        mlp.bias(l).write(B[l])

    bx = dev.wrap(x_dev.data_ptr(), x_dev.numel()).no_grad()
    by = dev.wrap(y_dev.data_ptr(), y_dev.numel()).no_grad()
    bl = dev.wrap(labels_dev.data_ptr(), batch, np.int32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def resident_step(metrics: bool):
        return mlp.step(bx, by, bl, batch, LR, grad_rows=global_batch, want_metrics=metrics)

    # ---------------- self-check at the headline configuration, before anything is timed
    # (a) the default arithmetic (3xFP16 tensor-core gemms, fused epilogues) against SL_GEMM_SIMT from identical weights: the SIMT
    #     gemm is the kernel the parity tests hold BIT-IDENTICAL to the oracle's restatement (tests/test_gpu_gemm.py), so this ties
    #     the benchmarked step to the oracle at the full 65536 x 4096 x 4096 / K = 65536 shapes;
    # (b) under data parallelism: the all-reduced bucket of the N ranks against the gradient one GPU computes on the whole global
    #     batch (rel 1e-5); after the timed steps, the parameters of all ranks must be bit-identical.
    parity = None
    if not args.no_parity_check:
        def grads(mode):
            dev.set_gemm_mode(mode)
            l, c_ = mlp.forward_backward(bx, by, bl, batch, grad_rows=global_batch)
            mlp.allreduce_grads()
            return l, c_, mlp.grad_bucket().read()
        l_ref, c_ref, g_ref = grads(S.GEMM_SIMT)
        l_got, c_got, g_got = grads({"tf32": S.GEMM_TF32, "3xtf32": S.GEMM_3XTF32, "3xf16": S.GEMM_3XF16}[args.gemm_mode])
        gmax = float(np.max(np.abs(g_ref)))
        gdiff = float(np.max(np.abs(g_got - g_ref)))
        tol_g = 1e-4 if args.gemm_mode != "tf32" else 5e-2
        parity = dict(reference="SL_GEMM_SIMT step (CUDA-core fp32, bit-identical to the oracle's gemm restatement) from the same weights and batch",
                      loss_sum_ref=l_ref, loss_sum=l_got, loss_rel_diff=abs(l_got - l_ref) / max(abs(l_ref), 1e-30),
                      correct_ref=int(c_ref), correct=int(c_got), grad_entries_compared=int(g_ref.size), grad_max_abs=gmax,
                      grad_max_abs_diff=gdiff, grad_rel_diff=gdiff / max(gmax, 1e-30), tol_grad_rel=tol_g, tol_loss_rel=1e-5 if args.gemm_mode != "tf32" else 1e-2)
        parity["ok"] = bool(parity["loss_rel_diff"] <= parity["tol_loss_rel"] and parity["grad_rel_diff"] <= tol_g and
                            abs(c_got - c_ref) <= max(2, batch // 1000) and np.all(np.isfinite(g_got)))
        if world > 1:
            if rank == 0:   # one GPU, whole global batch, a second device WITHOUT a communicator (no collective on this path)
                dev1 = CUDA(local_rank, cached=True, stream=stream.cuda_stream)
                dev1.set_gemm_mode({"tf32": S.GEMM_TF32, "3xtf32": S.GEMM_3XTF32, "3xf16": S.GEMM_3XF16}[args.gemm_mode])
                m1 = Mlp(dev1, DIMS, 0)
                m1.set_fused(not args.unfused)
                for l in range(len(DIMS) - 1):
                    m1.weights(l).write(W[l]); m1.bias(l).write(B[l])
                xs, ys, ls = [], [], []
                for r in range(world):
                    gr = torch.Generator(device="cuda").manual_seed(7 + r)
                    xs.append(torch.empty(batch, DIMS[0], device="cuda").uniform_(0, 1, generator=gr))
                    lr_ = torch.randint(0, DIMS[-1], (batch,), device="cuda", generator=gr, dtype=torch.int32)
                    yr = torch.zeros(batch, DIMS[-1], device="cuda"); yr[torch.arange(batch, device="cuda"), lr_.long()] = 1.0
                    ys.append(yr); ls.append(lr_)
                xg, yg, lg = torch.cat(xs), torch.cat(ys), torch.cat(ls)
                del xs, ys, ls
                torch.cuda.synchronize()
                l1, c1 = m1.forward_backward(dev1.wrap(xg.data_ptr(), xg.numel()).no_grad(), dev1.wrap(yg.data_ptr(), yg.numel()).no_grad(),
                                             dev1.wrap(lg.data_ptr(), global_batch, np.int32), global_batch, grad_rows=global_batch)
                g1 = m1.grad_bucket().read()
                d1 = float(np.max(np.abs(g_got - g1)))
                parity["dp"] = dict(ranks=world, one_gpu_global_batch_vs_allreduced_bucket_rel=d1 / max(float(np.max(np.abs(g1))), 1e-30), tol=1e-5)
                parity["ok"] = bool(parity["ok"] and parity["dp"]["one_gpu_global_batch_vs_allreduced_bucket_rel"] <= 1e-5)
                del m1, xg, yg, lg
                dev1.close()
                torch.cuda.empty_cache()
        for l in range(len(DIMS) - 1):   # forward_backward does not touch the parameters; rewrite them anyway: the timed run starts from W
            mlp.weights(l).write(W[l]); mlp.bias(l).write(B[l])

    # ---------------- device-resident leg (`value`)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    first_metrics = resident_step(True)
    for _ in range(max(args.warmup, 3) - 1):
        resident_step(True)
    barrier()
    sampler.mark_begin()
    launches0 = dev.launches
    capi.check(ctx, lib.sl_ctx_profile_begin(ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        resident_step(False)   # loss / accuracy are still computed on the device every step; only the host read is outside `value`
    mlp.flush()
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.mark_end()
    ms_total = e0.elapsed_time(e1)
    n_l, t_ms, t_fl = C.c_uint64(0), C.c_double(0), C.c_double(0)
    capi.check(ctx, lib.sl_ctx_profile_end(ctx, C.byref(n_l), C.byref(t_ms), C.byref(t_fl)))
    launches = dev.launches - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_total / args.steps)
    value = global_batch / (ms_step * 1e-3)
    loss_sum, correct = resident_step(True)
    first_loss = first_metrics[0] / batch

    # ---------------- end-to-end leg: host buffers in, loss/accuracy out, every step
    # two staging sets: batch i+1 uploads on the copy stream (sl_write_prefetch) while step i computes; every step's inputs are
    # copied exactly once, inside the timed region, and every step's loss / accuracy is read back (a blocking 8-byte read)
    stage = []
    for _ in range(2):
        sx, sy, sl_ = dev.zeros(batch * DIMS[0]), dev.zeros(batch * DIMS[-1]), dev.zeros(batch, np.int32)
        sx.no_grad(); sy.no_grad()
        stage.append((sx, sy, sl_))
    h2d = x_pin.numel() * 4 + y_pin.numel() * 4 + l_pin.numel() * 4
    d2h = 8

    def upload(slot):
        sx, sy, sl_ = stage[slot]
        capi.check(ctx, lib.sl_prefetch_release(ctx))   # the copy may not overtake compute that still reads this slot
        capi.check(ctx, lib.sl_write_prefetch(ctx, sx.ptr, x_pin.data_ptr(), x_pin.numel() * 4))
        capi.check(ctx, lib.sl_write_prefetch(ctx, sy.ptr, y_pin.data_ptr(), y_pin.numel() * 4))
        capi.check(ctx, lib.sl_write_prefetch(ctx, sl_.ptr, l_pin.data_ptr(), l_pin.numel() * 4))

    met_pin = torch.zeros(2 * (args.steps + 4), dtype=torch.float32, pin_memory=True)   # [loss_sum f32 | correct i32] per step
    met_ptr = mlp.metrics_ptr

    def e2e_run(nsteps):
        upload(0)
        for i in range(nsteps):
            capi.check(ctx, lib.sl_prefetch_wait(ctx))  # compute stream waits for batch i
            if i + 1 < nsteps:
                upload((i + 1) & 1)                     # batch i+1 goes up while step i runs
            sx, sy, sl_ = stage[i & 1]
            mlp.step(sx, sy, sl_, batch, LR, grad_rows=global_batch, want_metrics=False)
            # every step's loss / accuracy goes to (pinned) host memory, stream-ordered, without stalling the next launch
            capi.check(ctx, lib.sl_read_async(ctx, met_pin.data_ptr() + 8 * i, met_ptr, 8))

    e2e_run(2)
    barrier()
    e0.record(stream)
    e2e_run(args.steps)
    mlp.flush()
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    e2e_losses = met_pin[0:2 * args.steps:2].clone()
    assert bool(torch.isfinite(e2e_losses).all()) and float(e2e_losses.min()) > 0, "e2e leg: a step's loss did not arrive on the host"
    barrier()

    if args.breakdown:
        # in-situ per-kernel times (CUDA events around EVERY launch of a few extra steps; not part of any reported number).
        # EVERY rank runs the extra steps (the step contains the gradient all-reduce); rank 0 writes the table.
        buf = C.create_string_buffer(1 << 16)
        capi.check(ctx, lib.sl_ctx_profile_report(ctx, None, 0))     # switch on per-launch timing
        capi.check(ctx, lib.sl_ctx_profile_begin(ctx))
        bsteps = 5
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for _ in range(bsteps):
            resident_step(False)
        mlp.flush()
        b1.record(stream)
        torch.cuda.synchronize()
        capi.check(ctx, lib.sl_ctx_profile_report(ctx, buf, len(buf)))
        capi.check(ctx, lib.sl_ctx_profile_end(ctx, None, None, None))
        wall = b0.elapsed_time(b1) / bsteps
        rows = [r.rsplit(",", 2) for r in buf.value.decode().strip().splitlines()]
        tot = sum(float(r[2]) for r in rows) / bsteps
        barrier()
        with open(args.breakdown if rank == 0 else os.devnull, "w") as f:
            f.write(f"# per-kernel CUDA-event times inside the training step, N={world}, batch {batch}, {args.gemm_mode}; {bsteps} steps averaged\n")
            f.write(f"# step (events around every launch add gaps): {wall:.3f} ms; sum of kernels: {tot:.3f} ms\n")
            f.write(f"# {'kernel':60s} launches/step   ms/step   share\n")
            for name, n, ms in rows:
                f.write(f"{name[:62]:62s} {int(n) / bsteps:8.1f} {float(ms) / bsteps:10.4f} {100 * float(ms) / bsteps / tot:7.2f}%\n")
    if world > 1 and parity is not None:   # replicas bit-identical across ranks after all the steps above (same summed gradients everywhere)
        import zlib
        crc = zlib.crc32(mlp.params().read().tobytes())
        t = torch.tensor([crc], device="cuda", dtype=torch.int64)
        allc = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allc, t)
        same = all(int(a.item()) == crc for a in allc)
        if rank == 0:
            parity["dp"]["replica_param_crc32_identical_across_ranks"] = bool(same)
            parity["ok"] = bool(parity["ok"] and same)
    sweeps_out = None
    if rank == 0 and world == 1 and not args.no_sweeps:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sweeps as SW
        del mlp   # free the step's activations before the 16384^2 sweeps
        mlp = None
        stage.clear()
        torch.cuda.empty_cache()
        try:
            sweeps_out = SW.run_all(local_rank, args.sweep_budget)
        except Exception as e:   # a sweep failure must not lose the headline line
            sweeps_out = dict(error=repr(e))
    if rank == 0:
        pk = peaks()
        f16 = args.gemm_mode == "3xf16"
        # the pipe the kernel runs on: kind::f16 at the measured bf16 rate, kind::tf32 at half of it
        tf32_peak = pk["bf16_sustained"] if f16 else pk["bf16_sustained"] / 2.0
        eff = (t_fl.value / (t_ms.value * 1e-3)) / 1e12 if t_ms.value > 0 else 0.0
        mult = 1 if args.gemm_mode == "tf32" else 3
        roofline = dict(bound="tensor", kernel="gemm_tf32_2cta_kernel (tcgen05 cta_group::2 %s, TMA, TMEM)" % ("kind::f16" if f16 else "kind::tf32"),
                        achieved=eff, peak=tf32_peak, unit="TFLOP/s",
                        frac=eff / tf32_peak, traffic=(ncu_traffic() or {}).get("bytes_per_launch"), traffic_detail=ncu_traffic(),
                        peak_src=(f"{pk['src']}: bf16_tflops_sustained (kind::f16 runs at the bf16 rate); kernel timed inside a long step" if f16 else
                                  f"{pk['src']}: bf16_tflops_sustained/2 (dense TF32 = half the bf16 rate); kernel timed inside a long step"),
                        issued_tflops=eff * mult, pipe_util=eff * mult / tf32_peak,
                        # an f32 gemm's natural yardstick: the dense TF32 rate (half the measured bf16 rate)
                        frac_of_tf32_peak=eff / (pk["bf16_sustained"] / 2.0),
                        note=f"achieved = algorithmic 2MNK flops of the {n_l.value} tensor-core gemm launches / their summed CUDA-event time "
                             f"({t_ms.value / max(n_l.value, 1):.3f} ms avg); {args.gemm_mode} issues {mult}x those flops to the tensor pipe",
                        gemm_share_of_step=t_ms.value / ms_total if ms_total > 0 else None,
                        step_effective_tflops=step_flops(batch) / (ms_total / args.steps * 1e-3) / 1e12)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_step,
                    higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=WORKLOAD if not args.debug_per_gpu_batch else "DIAGNOSIS RUN (per-GPU batch overridden): " + WORKLOAD,
                                global_batch=global_batch, per_gpu_batch=batch, parallelism=f"dp{world}",
                                gemm_mode=args.gemm_mode, fused_epilogues=not args.unfused, dp_pipelined_update=bool(deferred), host_numa_node=numa, l2="inputs (>= 128 MiB per operand at every N) exceed the 126 MB L2; no flush needed"),
                    e2e=dict(value=global_batch / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, ms_per_step=e2e_ms),
                    gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                    training=dict(first_step_mean_loss=first_loss, last_step_mean_loss=loss_sum / batch, last_step_accuracy=correct / batch,
                                  init="W ~ U(-a, a), a = min(0.1, sqrt(6/(fan_in+fan_out))); b = 0"))
        if parity is not None:
            line["parity_check"] = parity
        if sweeps_out is not None:
            line["sweeps"] = sweeps_out
        if world == 1 and not args.no_cpu_baseline:
            cb, _ = cpu_reference_rate(1, 0, args.cpu_sample)   # one full-batch step: ~15-20 s of host time
            cb["sine_net_us_per_step"] = cpu_sine_net_us()      # the same CPU port beside sweeps.sine_net
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            print("PARITY CHECK FAILED: " + json.dumps(parity), file=sys.stderr, flush=True)
            rc_fail = True
    del mlp
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rc_fail:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--debug-per-gpu-batch", type=int, default=0, help="diagnosis only: override the per-GPU batch (not the BASELINE workload)")
    ap.add_argument("--breakdown", default=None, help="write an in-situ per-kernel time table of the step to this path")
    ap.add_argument("--gemm-mode", choices=["3xtf32", "tf32", "3xf16"], default="3xf16")
    ap.add_argument("--cpu-sample", type=int, default=GLOBAL_BATCH, help="batch of the CPU legs (default: the full global batch; smaller = a bounded sample)")
    ap.add_argument("--no-sweeps", action="store_true", help="skip the configs[0..3] sweeps appended to the N=1 line")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--sweep-budget", type=float, default=60.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="run the op-by-op tape instead of the fused-epilogue step (bit-identical results)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
