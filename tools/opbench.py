"""
Kernel-level sweep (BASELINE.json configs[1] and configs[2]): every op + grad at 16384 x 16384 f32 (HBM GB/s vs the
measured copy peak) and the gemm sweep (effective TFLOP/s, tensor-pipe utilisation = issued MMA flops / peak).
Timing: CUDA events on the stream the kernels are launched on, W warm-up + K timed launches, inputs >= 1 GiB (>> 126 MB L2).

    python tools/opbench.py [--ops] [--gemm] [--size 16384] [--reps 20] [--json out.json]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import sliced_b200 as S
from sliced_b200.raw import DeviceArray


def peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


class Bench:
    def __init__(self):
        torch.cuda.set_device(0)
        self.stream = torch.cuda.current_stream()
        self.ctx = S.Context(0, stream=self.stream.cuda_stream)
        self.keep = []

    def buf(self, n, fill="rand", dtype=torch.float32, lo=-1.0, hi=1.0):
        if fill == "rand":
            t = torch.empty(n, dtype=dtype, device="cuda").uniform_(lo, hi)
        elif fill == "zeros":
            t = torch.zeros(n, dtype=dtype, device="cuda")
        else:
            t = torch.full((n,), float(fill), dtype=dtype, device="cuda")
        self.keep.append(t)
        return DeviceArray(self.ctx, n, np.float32 if dtype == torch.float32 else np.int32, ptr=t.data_ptr(), owner=t)

    def time(self, fn, reps, warmup=3):
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


def run_ops(B, R, C, reps, out):
    ctx = B.ctx
    N = R * C
    hbm, _, src = peaks()
    x, y, z = B.buf(N), B.buf(N), B.buf(N)
    g1, g2 = B.buf(N), B.buf(N)
    pos = B.buf(N, lo=0.5, hi=2.0)
    rv, cv = B.buf(C), B.buf(R)
    ro, co = B.buf(C), B.buf(R)
    idx = DeviceArray(ctx, R, np.int32)
    L = ctx.lib
    h = ctx.h
    F = S.F32
    ops = [
        ("add fwd", 12, lambda: L.sl_binary_ew(h, F, S.ADD, x.ptr, y.ptr, z.ptr, N)),
        ("mul fwd", 12, lambda: L.sl_binary_ew(h, F, S.MUL, x.ptr, y.ptr, z.ptr, N)),
        ("add grad", 20, lambda: L.sl_binary_ew_grad(h, F, S.ADD, x.ptr, y.ptr, g1.ptr, g2.ptr, z.ptr, N)),
        ("mul grad", 28, lambda: L.sl_binary_ew_grad(h, F, S.MUL, x.ptr, y.ptr, g1.ptr, g2.ptr, z.ptr, N)),
        ("square fwd", 8, lambda: L.sl_unary(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, z.ptr, N)),
        ("square grad", 16, lambda: L.sl_unary_grad(h, F, S.UN_SQUARE, 0.0, 0.0, x.ptr, g1.ptr, z.ptr, N)),
        ("pow3 fwd", 8, lambda: L.sl_unary(h, F, S.UN_POW, 3.0, 0.0, pos.ptr, z.ptr, N)),
        ("pow3 grad", 16, lambda: L.sl_unary_grad(h, F, S.UN_POW, 3.0, 0.0, pos.ptr, g1.ptr, z.ptr, N)),
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
    # make max_cols/max_rows outputs consistent for the grads
    L.sl_max_rows(h, F, R, C, x.ptr, ro.ptr, None)
    L.sl_max_cols(h, F, R, C, x.ptr, co.ptr, idx.ptr)
    print(f"{'op':30s} {'B/elem':>6s} {'ms':>8s} {'GB/s':>8s} {'frac of ' + src + ' ' + str(hbm):>24s}")
    for name, bpe, fn in ops:
        rc = fn()
        assert rc == 0, (name, ctx.lib.sl_last_error_string(ctx.h))
        ms, best = B.time(fn, reps)
        gbs = bpe * N / (ms * 1e-3) / 1e9
        print(f"{name:30s} {bpe:6d} {ms:8.3f} {gbs:8.0f} {gbs / hbm:24.3f}", flush=True)
        out.append(dict(kind="op", op=name, rows=R, cols=C, bytes_per_elem=bpe, ms=ms, ms_best=best, gbs=gbs, frac=gbs / hbm, peak=hbm, peak_src=src))


def run_gemm(B, sizes, reps, out, cfgs=("0",)):
    ctx = B.ctx
    _, bf16, src = peaks()
    tf32_peak = bf16 / 2
    print(f"{'gemm':28s} {'mode':7s} {'cfg':3s} {'ms':>9s} {'eff TF/s':>9s} {'issued TF/s':>11s} {'pipe util (of ' + src + ' bf16 ' + str(bf16) + ' / tf32 ' + str(tf32_peak) + ')':>36s}")
    for (m, n, k) in sizes:
        a, b = B.buf(m * k), B.buf(k * n)
        c = B.buf(m * n, "zeros")
        # pipe util: issued flops against the rate of the pipe the mode runs on (kind::f16 = the bf16 rate, kind::tf32 = half of it)
        for mode, name, mult, pipe_peak in ((S.GEMM_3XF16, "3xf16", 3, bf16), (S.GEMM_3XTF32, "3xtf32", 3, tf32_peak), (S.GEMM_TF32, "tf32", 1, tf32_peak)):
            for cfg in cfgs:
                os.environ["SLICED_GEMM_CFG"] = cfg
                for (ta, tb, tag) in ((0, 0, "NN"), (0, 1, "NT"), (1, 0, "TN")):
                    fn = lambda: ctx.lib.sl_gemm_ex(ctx.h, S.F32, ta, tb, m, n, k, a.ptr, b.ptr, c.ptr, 0, mode)
                    rc = fn()
                    assert rc == 0, ctx.lib.sl_last_error_string(ctx.h)
                    ms, best = B.time(fn, reps)
                    eff = 2.0 * m * n * k / (ms * 1e-3) / 1e12
                    print(f"{tag} {m}x{n}x{k:<14d} {name:7s} {cfg:3s} {ms:9.3f} {eff:9.1f} {eff * mult:11.1f} {eff * mult / pipe_peak:36.3f}", flush=True)
                    out.append(dict(kind="gemm", layout=tag, m=m, n=n, k=k, mode=name, cfg=cfg, ms=ms, ms_best=best, eff_tflops=eff,
                                    issued_tflops=eff * mult, pipe_util=eff * mult / pipe_peak, pipe_peak=pipe_peak, peak_tf32=tf32_peak, peak_src=src))
        os.environ.pop("SLICED_GEMM_CFG", None)
        B.keep.clear()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", action="store_true")
    ap.add_argument("--gemm", action="store_true")
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--gemm-sizes", type=str, default="2048,4096,8192")
    ap.add_argument("--cfgs", type=str, default="0")
    ap.add_argument("--json", type=str, default="")
    args = ap.parse_args()
    B = Bench()
    out = []
    if args.ops or not args.gemm:
        run_ops(B, args.size, args.size, args.reps, out)
        B.keep.clear(); torch.cuda.empty_cache()
    if args.gemm:
        sizes = [(int(s),) * 3 for s in args.gemm_sizes.split(",")]
        run_gemm(B, sizes, max(3, args.reps // 2), out, cfgs=tuple(args.cfgs.split(",")))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)
