"""
Sweeps of BASELINE.json configs[0..3], callable from bench.py (the `sweeps` key of the N=1 line) and from tools/opbench.py:

  ops      configs[1]  every op + grad at 16384 x 16384 f32: ms, GB/s, fraction of the MEASURED copy peak and of the 8 TB/s spec
  gemm     configs[2]  M=N=K 512..16384 x {3xf16, 3xtf32, tf32} x {NN, NT, TN}: the WHOLE call (operand prep + MMA kernel + split-K
                       fold) -> effective TFLOP/s and tensor-pipe utilisation (issued MMA flops / peak of the pipe the mode runs on)
  chained  configs[3]  chained_perf.rs graph at 2^28: one fused pass vs the 5 + 5 separate launches of the unfused tape
  sine_net configs[0]  sine_net.rs at the shipped sizes: us/step eager tape vs CUDA-graph replay, launches/step

Timing: CUDA events on the stream the kernels are launched on, warm-up + median of the timed launches; the big operands are
>= 1 GiB (>> 126 MB L2) and the small gemm sizes rotate between several operand sets so that they do not sit in L2.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HBM_SPEC_GBS = 8000.0   # BASELINE.json north_star: "~8 TB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class Bench:
    """torch owns the memory and the stream; the library runs on that stream through its C ABI"""

    def __init__(self, device_index=0):
        import torch

        import sliced_b200 as S
        self.torch, self.S = torch, S
        torch.cuda.set_device(device_index)
        self.stream = torch.cuda.current_stream()
        self.ctx = S.Context(device_index, stream=self.stream.cuda_stream)
        self.keep = []

    def buf(self, n, fill="rand", dtype=None, lo=-1.0, hi=1.0):
        torch = self.torch
        from sliced_b200.raw import DeviceArray
        dtype = dtype or torch.float32
        if fill == "rand":
            t = torch.empty(n, dtype=dtype, device="cuda").uniform_(lo, hi)
        elif fill == "zeros":
            t = torch.zeros(n, dtype=dtype, device="cuda")
        else:
            t = torch.full((n,), float(fill), dtype=dtype, device="cuda")
        self.keep.append(t)
        return DeviceArray(self.ctx, n, np.float32 if dtype == torch.float32 else np.int32, ptr=t.data_ptr(), owner=t)

    def release(self):
        self.keep.clear()
        self.torch.cuda.empty_cache()

    def time(self, fn, reps, warmup=3):
        torch = self.torch
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(self.stream)
            fn()
            b.record(self.stream)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return ts[len(ts) // 2], ts[0]

    def close(self):
        self.release()
        self.ctx.close()


def op_table(B: Bench, R: int, C: int):
    """(name, algorithmic bytes per element [SURVEY 8d], launcher) for every op + grad of config 2"""
    S, ctx = B.S, B.ctx
    from sliced_b200.raw import DeviceArray
    N = R * C
    x, y, z = B.buf(N), B.buf(N), B.buf(N)
    g1, g2 = B.buf(N), B.buf(N)
    pos = B.buf(N, lo=0.5, hi=2.0)
    rv, cv = B.buf(C), B.buf(R)
    ro, co = B.buf(C), B.buf(R)
    idx = DeviceArray(ctx, R, np.int32)
    B.keep.append(idx)
    L, h, F = ctx.lib, ctx.h, S.F32
    ops = [
        ("add fwd", 12, lambda: L.sl_binary_ew(h, F, S.ADD, x.ptr, y.ptr, z.ptr, N)),
        ("mul fwd", 12, lambda: L.sl_binary_ew(h, F, S.MUL, x.ptr, y.ptr, z.ptr, N)),
        ("add grad", 20, lambda: L.sl_binary_ew_grad(h, F, S.ADD, x.ptr, y.ptr, g1.ptr, g2.ptr, z.ptr, N)),
        ("mul grad", 28, lambda: L.sl_binary_ew_grad(h, F, S.MUL, x.ptr, y.ptr, g1.ptr, g2.ptr, z.ptr, N)),
        ("square fwd", 8, lambda: L.sl_unary(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, z.ptr, N)),
        ("square grad", 16, lambda: L.sl_unary_grad(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, g1.ptr, z.ptr, N)),
        ("pow3 fwd", 8, lambda: L.sl_unary(h, F, S.UN_POW, 3.0, 0.0, pos.ptr, z.ptr, N)),
        ("pow3 grad", 16, lambda: L.sl_unary_grad(h, F, S.UN_POW, 3.0, 0.0, pos.ptr, g1.ptr, z.ptr, N)),
        ("pow2.5 fwd (powf)", 8, lambda: L.sl_unary(h, F, S.UN_POW, 2.5, 0.0, pos.ptr, z.ptr, N)),
        ("relu fwd", 8, lambda: L.sl_unary(h, F, S.UN_RELU, 0.0, 0.0, x.ptr, z.ptr, N)),
        ("relu grad", 16, lambda: L.sl_unary_grad(h, F, S.UN_RELU, 0.0, 0.0, x.ptr, g1.ptr, z.ptr, N)),
        ("tanh fwd", 8, lambda: L.sl_unary(h, F, S.UN_TANH, 0.0, 0.0, x.ptr, z.ptr, N)),
        ("add_row fwd", 8, lambda: L.sl_add_row(h, F, R, C, x.ptr, rv.ptr, z.ptr)),
        ("add_row_mut", 8, lambda: L.sl_add_row_mut(h, F, R, C, g1.ptr, rv.ptr)),
        ("add_row grad", 8, lambda: L.sl_add_row_grad(h, F, R, C, g1.ptr, ro.ptr, x.ptr)),
        ("add_row_mut grad", 4, lambda: L.sl_add_row_mut_grad(h, F, R, C, ro.ptr, x.ptr)),
        ("sum_rows fwd", 4, lambda: L.sl_sum_rows(h, F, R, C, x.ptr, ro.ptr)),
        ("sum_cols fwd", 4, lambda: L.sl_sum_cols(h, F, R, C, x.ptr, co.ptr)),
        ("mean_rows fwd", 4, lambda: L.sl_mean_rows(h, F, R, C, x.ptr, ro.ptr)),
        ("mean_cols fwd", 4, lambda: L.sl_mean_cols(h, F, R, C, x.ptr, co.ptr)),
        ("max_rows fwd", 4, lambda: L.sl_max_rows(h, F, R, C, x.ptr, ro.ptr, None)),
        ("max_cols fwd(+idx)", 4, lambda: L.sl_max_cols(h, F, R, C, x.ptr, co.ptr, idx.ptr)),
        ("sum_rows grad", 8, lambda: L.sl_sum_rows_grad(h, F, R, C, g1.ptr, rv.ptr)),
        ("sum_cols grad", 8, lambda: L.sl_sum_cols_grad(h, F, R, C, g1.ptr, cv.ptr)),
        ("mean_rows grad", 8, lambda: L.sl_mean_rows_grad(h, F, R, C, g1.ptr, rv.ptr)),
        ("mean_cols grad", 8, lambda: L.sl_mean_cols_grad(h, F, R, C, g1.ptr, cv.ptr)),
        ("max_rows grad", 12, lambda: L.sl_max_rows_grad(h, F, R, C, ro.ptr, x.ptr, g1.ptr, rv.ptr)),
        ("max_cols grad(scan)", 4, lambda: L.sl_max_cols_grad(h, F, R, C, co.ptr, x.ptr, g1.ptr, cv.ptr)),
        ("softmax fwd", 8, lambda: L.sl_softmax(h, F, R, C, x.ptr, z.ptr)),
        ("softmax grad", 12, lambda: L.sl_softmax_grad(h, F, R, C, g1.ptr, z.ptr, y.ptr)),
        ("transpose", 8, lambda: L.sl_transpose(h, F, R, C, x.ptr, z.ptr, 0)),
        ("transpose acc", 12, lambda: L.sl_transpose(h, F, R, C, x.ptr, g1.ptr, 1)),
        ("sgd step", 12, lambda: L.sl_sgd_step(h, F, g1.ptr, x.ptr, 0.1, N)),
        ("chained fwd (fused 5 ops)", 12, lambda: L.sl_chained_fwd(h, F, x.ptr, y.ptr, z.ptr, N)),
        ("chained bwd (fused tape)", 28, lambda: L.sl_chained_bwd(h, F, x.ptr, y.ptr, g1.ptr, g2.ptr, z.ptr, N)),
    ]
    # make max_cols / max_rows outputs consistent for the grads
    L.sl_max_rows(h, F, R, C, x.ptr, ro.ptr, None)
    L.sl_max_cols(h, F, R, C, x.ptr, co.ptr, idx.ptr)
    return ops, N


def run_ops(B: Bench, R=16384, C=16384, reps=10, verbose=False):
    pk = peaks()
    ops, N = op_table(B, R, C)
    rows = []
    if verbose:
        print(f"{'op':30s} {'B/elem':>6s} {'ms':>8s} {'GB/s':>8s} {'of ' + pk['src'] + ' ' + str(pk['hbm']):>22s} {'of spec 8000':>13s}")
    for name, bpe, fn in ops:
        rc = fn()
        assert rc == 0, (name, B.ctx.lib.sl_last_error_string(B.ctx.h))
        ms, best = B.time(fn, reps)
        gbs = bpe * N / (ms * 1e-3) / 1e9
        rows.append(dict(op=name, bytes_per_elem=bpe, ms=round(ms, 4), gbs=round(gbs, 1), frac_measured=round(gbs / pk["hbm"], 4),
                         frac_spec=round(gbs / HBM_SPEC_GBS, 4)))
        if verbose:
            print(f"{name:30s} {bpe:6d} {ms:8.3f} {gbs:8.0f} {gbs / pk['hbm']:22.3f} {gbs / HBM_SPEC_GBS:13.3f}", flush=True)
    B.release()
    return dict(rows_cols=[R, C], dtype="f32", reps=reps, peak_measured_gbs=pk["hbm"], peak_spec_gbs=HBM_SPEC_GBS, peak_src=pk["src"], ops=rows)


GEMM_MODES = (("3xf16", 3, 3, "bf16"), ("3xtf32", 0, 3, "tf32"), ("tf32", 1, 1, "tf32"))


def run_gemm(B: Bench, sizes=(512, 1024, 2048, 4096, 8192, 16384), reps=5, verbose=False, modes=None):
    """whole-call timing (prep + MMA + fold).  pipe peak: kind::f16 at the measured bf16 burst rate, kind::tf32 at half of it."""
    S, ctx = B.S, B.ctx
    pk = peaks()
    rows = []
    if verbose:
        print(f"{'gemm':24s} {'mode':7s} {'ms':>9s} {'eff TF/s':>9s} {'issued':>8s} {'pipe util':>10s}  (pipe peak: bf16 {pk['bf16']} / tf32 {pk['bf16'] / 2} {pk['src']})")
    for n in sizes:
        # small problems: rotate operand sets so that successive launches do not find their operands in the 126 MB L2
        nsets = max(1, min(8, int(200e6 // (3 * n * n * 4)) + 1)) if n <= 2048 else 1
        sets = [(B.buf(n * n), B.buf(n * n), B.buf(n * n, "zeros")) for _ in range(nsets)]
        for name, code, mult, pipe in GEMM_MODES:
            if modes and name not in modes:
                continue
            pipe_peak = pk["bf16"] if pipe == "bf16" else pk["bf16"] / 2
            for ta, tb, tag in ((0, 0, "NN"), (0, 1, "NT"), (1, 0, "TN")):
                it = [0]

                def fn():
                    a, b, c = sets[it[0] % nsets]
                    it[0] += 1
                    return ctx.lib.sl_gemm_ex(ctx.h, S.F32, ta, tb, n, n, n, a.ptr, b.ptr, c.ptr, 0, code)
                rc = fn()
                assert rc == 0, ctx.lib.sl_last_error_string(ctx.h)
                r = reps if n >= 4096 else reps * 4
                ms, best = B.time(fn, r, warmup=max(2, nsets))
                eff = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
                rows.append(dict(layout=tag, n=n, mode=name, ms=round(ms, 4), eff_tflops=round(eff, 1), issued_tflops=round(eff * mult, 1),
                                 pipe_util=round(eff * mult / pipe_peak, 4)))
                if verbose:
                    print(f"{tag} {n:<6d}^3              {name:7s} {ms:9.3f} {eff:9.1f} {eff * mult:8.1f} {eff * mult / pipe_peak:10.3f}", flush=True)
        B.release()
    return dict(sizes=list(sizes), reps=reps, timed="whole sl_gemm_ex call: operand split + MMA kernel + split-K fold",
                pipe_peak_bf16=pk["bf16"], pipe_peak_tf32=pk["bf16"] / 2, peak_src=pk["src"], rows=rows)


def run_chained(B: Bench, log2n=28, reps=5):
    """examples/chained_perf.rs:86-114: out = x^2 * x + (b + x) * b and its tape, fused (1 + 1 launches) vs unfused (5 + 5)"""
    S, ctx = B.S, B.ctx
    L, h, F = ctx.lib, ctx.h, S.F32
    N = 1 << log2n
    x, b, out = B.buf(N, lo=-2, hi=2), B.buf(N, lo=-2, hi=2), B.buf(N)
    xg, bg, og = B.buf(N, "zeros"), B.buf(N, "zeros"), B.buf(N, 1.0)
    t = [B.buf(N) for _ in range(4)]          # squared, add, mul_b, mul
    tg = [B.buf(N, "zeros") for _ in range(4)]

    def fused_fwd(): return L.sl_chained_fwd(h, F, x.ptr, b.ptr, out.ptr, N)
    def fused_bwd(): return L.sl_chained_bwd(h, F, x.ptr, b.ptr, xg.ptr, bg.ptr, og.ptr, N)

    def unfused_fwd():
        L.sl_unary(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, t[0].ptr, N)
        L.sl_binary_ew(h, F, S.ADD, b.ptr, x.ptr, t[1].ptr, N)
        L.sl_binary_ew(h, F, S.MUL, t[1].ptr, b.ptr, t[2].ptr, N)
        L.sl_binary_ew(h, F, S.MUL, t[0].ptr, x.ptr, t[3].ptr, N)
        return L.sl_binary_ew(h, F, S.ADD, t[3].ptr, t[2].ptr, out.ptr, N)

    def unfused_bwd():   # the five grad closures in reverse registration order (ACC into the intermediates' gradients)
        L.sl_binary_ew_grad(h, F, S.ADD, t[3].ptr, t[2].ptr, tg[3].ptr, tg[2].ptr, og.ptr, N)
        L.sl_binary_ew_grad(h, F, S.MUL, t[0].ptr, x.ptr, tg[0].ptr, xg.ptr, tg[3].ptr, N)
        L.sl_binary_ew_grad(h, F, S.MUL, t[1].ptr, b.ptr, tg[1].ptr, bg.ptr, tg[2].ptr, N)
        L.sl_binary_ew_grad(h, F, S.ADD, b.ptr, x.ptr, bg.ptr, xg.ptr, tg[1].ptr, N)
        return L.sl_unary_grad(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, xg.ptr, tg[0].ptr, N)
    res = {}
    for name, fn, bpe in (("fused_fwd", fused_fwd, 12), ("fused_bwd", fused_bwd, 28), ("unfused_fwd", unfused_fwd, 56), ("unfused_bwd", unfused_bwd, 112)):
        assert fn() == 0
        ms, _ = B.time(fn, reps)
        res[name] = dict(ms=round(ms, 3), bytes_per_elem=bpe, gbs=round(bpe * N / (ms * 1e-3) / 1e9, 1))
    res["n"] = N
    res["speedup_fwd"] = round(res["unfused_fwd"]["ms"] / res["fused_fwd"]["ms"], 2)
    res["speedup_bwd"] = round(res["unfused_bwd"]["ms"] / res["fused_bwd"]["ms"], 2)
    B.release()
    res["tape_n100"] = run_tape_overhead()
    return res


def run_tape_overhead(n=100, iters=300):
    """tape overhead at N = 100 (cf. tests/test_combination.rs:103-105): the same graph through the host tape
    (5 forward ops + their 5 grad closures), wall clock per op, on the plain device (one launch per op) and on the fusing device (the
    chain is recorded and runs as one forward + one backward launch)."""
    from sliced_b200.host import CUDA
    out = {}
    xs, bs = np.full(n, 1.3, np.float32), np.full(n, 2.1, np.float32)
    for name in ("plain", "fused"):
        dev = CUDA(0, cached=True)
        dev.set_fusion(name == "fused")
        x, b = dev.buffer(xs), dev.buffer(bs)

        def step():
            dev.zero_grad()
            o = dev.add(dev.mul(dev.square(x), x), dev.mul(dev.add(b, x), b))
            o.backward()
            return o
        for _ in dev.range(5):
            step()
        dev.sync()
        l0 = dev.launches
        t0 = time.perf_counter()
        for _ in dev.range(iters):    # `for _ in device.range(..)`: the Cached cursor rewinds every iteration
            step()
        dev.sync()
        dt = time.perf_counter() - t0
        out[name] = dict(us_per_graph=round(1e6 * dt / iters, 1), us_per_op=round(1e6 * dt / iters / 10, 2), launches_per_graph=round((dev.launches - l0) / iters, 1))
        del x, b
        dev.close()
    out["n"] = n
    out["ops_per_graph"] = 10
    return out


def run_sine_net(device_index=0, iters=300):
    """examples/sine_net.rs:119-166 at the shipped sizes: us/step and launches/step, eager tape vs CUDA-graph replay"""
    from sliced_b200.host import CUDA, Mlp
    dims = [1, 64, 64, 1]
    xs = (np.arange(1000) / 1000.0).astype(np.float32)
    ys = np.sin(2.0 * xs * np.float32(np.pi)).astype(np.float32)
    rng = np.random.default_rng(0)
    W0 = [rng.uniform(-0.5, 0.5, dims[i] * dims[i + 1]).astype(np.float32) for i in range(3)]
    out = {}
    for mode in ("eager", "graph", "one_launch"):   # one_launch: sl_mlp_small_step (the whole step in one cluster launch)
        dev = CUDA(device_index, cached=True)
        mlp = Mlp(dev, dims, 1)
        mlp.set_fused(mode == "one_launch")
        for l in range(3):
            mlp.weights(l).write(W0[l])
        dx, dy = dev.buffer(xs).no_grad(), dev.buffer(ys).no_grad()
        fn = mlp.step_replay if mode == "graph" else mlp.step
        for _ in range(3):
            fn(dx, dy, None, 1000, 1e-4)
        dev.sync()
        l0 = dev.launches
        t0 = time.perf_counter()
        for _ in range(iters):
            fn(dx, dy, None, 1000, 1e-4, want_metrics=False)
        dev.sync()
        dt = time.perf_counter() - t0
        loss = mlp.step(dx, dy, None, 1000, 1e-4)[0] / 1000
        out[mode] = dict(us_per_step=round(1e6 * dt / iters, 1), launches_per_step=round((dev.launches - l0) / iters, 1), mean_loss=loss)
        del mlp, dx, dy
        dev.close()
    out["iters"] = iters
    return out


def run_all(device_index=0, budget_s=60.0, verbose=False):
    """everything, inside a wall-clock budget (sections that would start past the budget are skipped and named)"""
    t0 = time.perf_counter()
    res, skipped = {}, []
    B = Bench(device_index)
    try:
        for key, fn in (("gemm", lambda: run_gemm(B, verbose=verbose)), ("ops", lambda: run_ops(B, verbose=verbose)),
                        ("chained", lambda: run_chained(B)), ("sine_net", lambda: run_sine_net(device_index))):
            if time.perf_counter() - t0 > budget_s:
                skipped.append(key)
                continue
            res[key] = fn()
    finally:
        B.close()
    res["wall_s"] = round(time.perf_counter() - t0, 1)
    if skipped:
        res["skipped_over_budget"] = skipped
    return res


def write_tables(sw: dict, prefix: str):
    """profiles/<prefix>_opbench_ops_16384.txt and _gemm_sweep.txt from a `sweeps` dict (bench.py line)"""
    if "ops" in sw:
        o = sw["ops"]
        with open(prefix + "_opbench_ops_16384.txt", "w") as f:
            f.write(f"# op sweep {o['rows_cols'][0]} x {o['rows_cols'][1]} f32, median of {o['reps']} launches (CUDA events); "
                    f"peaks: measured copy {o['peak_measured_gbs']} GB/s ({o['peak_src']}), spec {o['peak_spec_gbs']} GB/s\n")
            f.write(f"{'op':30s} {'B/elem':>6s} {'ms':>8s} {'GB/s':>8s} {'of measured':>12s} {'of spec':>8s}\n")
            for r in o["ops"]:
                f.write(f"{r['op']:30s} {r['bytes_per_elem']:6d} {r['ms']:8.3f} {r['gbs']:8.0f} {r['frac_measured']:12.3f} {r['frac_spec']:8.3f}\n")
    if "gemm" in sw:
        g = sw["gemm"]
        with open(prefix + "_opbench_gemm_sweep.txt", "w") as f:
            f.write(f"# gemm sweep M=N=K, f32 in / f32 out; timed: {g['timed']}; pipe peaks ({g['peak_src']}): kind::f16 {g['pipe_peak_bf16']} TF/s, "
                    f"kind::tf32 {g['pipe_peak_tf32']} TF/s; pipe util = issued MMA flops / pipe peak (3 MMAs per product in the 3x modes)\n")
            f.write(f"{'layout':6s} {'n':>6s} {'mode':7s} {'ms':>9s} {'eff TF/s':>9s} {'issued':>8s} {'pipe util':>10s}\n")
            for r in g["rows"]:
                f.write(f"{r['layout']:6s} {r['n']:6d} {r['mode']:7s} {r['ms']:9.3f} {r['eff_tflops']:9.1f} {r['issued_tflops']:8.1f} {r['pipe_util']:10.3f}\n")
    with open(prefix + "_sweeps_misc.json", "w") as f:
        json.dump({k: v for k, v in sw.items() if k not in ("ops", "gemm")}, f, indent=1)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-bench-json", default="", help="write the profiles/ tables from the `sweeps` key of a bench.py line")
    ap.add_argument("--prefix", default=os.path.join(ROOT, "profiles", "r2"))
    ap.add_argument("--budget", type=float, default=120.0)
    a = ap.parse_args()
    if a.from_bench_json:
        line = [ln for ln in open(a.from_bench_json).read().splitlines() if ln.startswith("{")][-1]
        write_tables(json.loads(line)["sweeps"], a.prefix)
    else:
        sw = run_all(0, a.budget, verbose=True)
        print(json.dumps({k: v for k, v in sw.items() if k not in ("ops", "gemm")}, indent=1))
        write_tables(sw, a.prefix)
