"""
Kernel-level sweep CLI (BASELINE.json configs[1] and configs[2]); the sweeps themselves live in tools/sweeps.py (bench.py runs the
same code for the `sweeps` key of its N=1 line).

    python tools/opbench.py [--ops] [--gemm] [--size 16384] [--reps 20] [--gemm-sizes 512,1024,2048,4096,8192,16384] [--modes 3xf16,tf32] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import sweeps

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", action="store_true")
    ap.add_argument("--gemm", action="store_true")
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--gemm-sizes", type=str, default="512,1024,2048,4096,8192,16384")
    ap.add_argument("--modes", type=str, default="")
    ap.add_argument("--json", type=str, default="")
    args = ap.parse_args()
    B = sweeps.Bench()
    out = {}
    if args.ops or not args.gemm:
        out["ops"] = sweeps.run_ops(B, args.size, args.size, args.reps, verbose=True)
    if args.gemm:
        out["gemm"] = sweeps.run_gemm(B, tuple(int(s) for s in args.gemm_sizes.split(",")), max(3, args.reps // 2), verbose=True,
                                      modes=set(args.modes.split(",")) if args.modes else None)
    B.close()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)
