"""One rank of the on-hardware data-parallel parity check (launched by tests/test_gpu_dp.py through torch.distributed.run, one
process per GPU, NCCL).  Every rank trains its batch shard with the fused device step; rank 0 additionally computes the same
step on ONE GPU over the whole batch (a second device without a communicator) and replays it with the CPU oracle.  Writes a JSON
report; asserts nothing itself."""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main(out_path, dims, batch, steps, lr):
    import oracle as O
    import sliced_b200 as S
    from sliced_b200 import capi, dp
    from sliced_b200.host import CUDA, Mlp
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(21)
    x = rng.uniform(0, 1, (batch, dims[0])).astype(np.float32)
    labels = rng.integers(0, dims[-1], batch).astype(np.int32)
    y = np.zeros((batch, dims[-1]), np.float32); y[np.arange(batch), labels] = 1
    a = [min(0.1, (6.0 / (dims[i] + dims[i + 1])) ** 0.5) for i in range(len(dims) - 1)]
    W = [(rng.uniform(-1, 1, dims[i] * dims[i + 1]) * a[i]).astype(np.float32) for i in range(len(dims) - 1)]
    B = [np.zeros(dims[i + 1], np.float32) for i in range(len(dims) - 1)]
    lo, hi = dp.shard_rows(batch, world, rank)

    def make(dev):
        m = Mlp(dev, dims, 0)
        m.set_fused(True)
        for l in range(len(dims) - 1):
            m.weights(l).write(W[l]); m.bias(l).write(B[l])
        return m
    dev = CUDA(local, cached=True)
    dp.init_comm(capi.load(), dev.ctx_handle, dist, rank, world)
    mlp = make(dev)
    mlp.set_deferred(os.environ.get("SLICED_TEST_DEFERRED", "0") == "1")   # cross-step pipelined join + SGD (Mlp::set_deferred)
    dx, dy, dl = dev.buffer(x[lo:hi]).no_grad(), dev.buffer(y[lo:hi]).no_grad(), dev.buffer(labels[lo:hi])
    # step 1 in two phases so that the summed bucket can be read before SGD consumes it
    l0, c0 = mlp.forward_backward(dx, dy, dl, hi - lo, grad_rows=batch)
    mlp.allreduce_grads()
    bucket = mlp.grad_bucket().read()
    mlp.sgd(lr)
    hist = [(l0, c0)] + [mlp.step(dx, dy, dl, hi - lo, lr, grad_rows=batch) for _ in range(steps - 1)]
    params = mlp.params().read()
    crc = zlib.crc32(params.tobytes())
    t = torch.tensor([crc], device="cuda", dtype=torch.int64)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    m = torch.tensor([[h[0], float(h[1])] for h in hist], device="cuda", dtype=torch.float64)
    dist.all_reduce(m)   # global loss sum / correct count per step
    rep = None
    if rank == 0:
        dev1 = CUDA(local, cached=True)   # no communicator: single-GPU reference of the same global step
        m1 = make(dev1)
        fx, fy, fl = dev1.buffer(x).no_grad(), dev1.buffer(y).no_grad(), dev1.buffer(labels)
        m1.forward_backward(fx, fy, fl, batch, grad_rows=batch)
        bucket1 = m1.grad_bucket().read()
        m1.sgd(lr)
        h1 = [m1.step(fx, fy, fl, batch, lr) for _ in range(steps - 1)]
        params1 = m1.params().read()
        O.use_openblas(os.cpu_count() or 1)
        Wo, Bo = [w.copy() for w in W], [b.copy() for b in B]
        seg, off = [], 0
        for i in range(len(dims) - 1):     # flat parameter layout of Mlp: segments padded to 64 floats
            for n in (dims[i] * dims[i + 1], dims[i + 1]):
                seg.append((off, n)); off += (n + 63) // 64 * 64

        def flat(ws, bs):
            f = np.zeros_like(params1)
            for (o, n), p in zip(seg, [p for pair in zip(ws, bs) for p in pair]):
                f[o:o + n] = p
            return f
        r0 = O.mlp_step(0, dims, x.ravel(), y.ravel(), labels, Wo, Bo, lr, want_grads=True)
        bucket_o = flat(r0[2], r0[3])
        ho = [r0[:2]] + [O.mlp_step(0, dims, x.ravel(), y.ravel(), labels, Wo, Bo, lr)[:2] for _ in range(steps - 1)]
        O.use_naive_gemm()
        flat_o = flat(Wo, Bo)
        gmax = float(np.max(np.abs(bucket1)))
        rep = dict(world=world, dims=dims, batch=batch, steps=steps,
                   bucket_rel_diff_vs_one_gpu=float(np.max(np.abs(bucket - bucket1))) / gmax,
                   bucket_rel_diff_vs_oracle=float(np.max(np.abs(bucket - bucket_o))) / gmax,
                   one_gpu_bucket_rel_diff_vs_oracle=float(np.max(np.abs(bucket1 - bucket_o))) / gmax,
                   params_rel_diff_vs_one_gpu=float(np.max(np.abs(params - params1))) / float(np.max(np.abs(params1))),
                   params_rel_diff_vs_oracle=float(np.max(np.abs(params - flat_o))) / float(np.max(np.abs(flat_o))),
                   crc_identical_across_ranks=all(int(c.item()) == crc for c in allc), params_crc=crc,
                   loss_dp=[float(v) for v in m[:, 0].cpu()], correct_dp=[int(v) for v in m[:, 1].cpu()],
                   loss_oracle=[float(h[0]) for h in ho], correct_oracle=[int(h[1]) for h in ho])
        del m1, fx, fy, fl
        dev1.close()
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(rep, f)
    del mlp, dx, dy, dl
    dist.destroy_process_group()
    dev.close()


if __name__ == "__main__":
    main(sys.argv[1], json.loads(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
