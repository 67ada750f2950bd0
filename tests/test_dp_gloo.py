"""
The N > 1 path on CPU: world_size 2 over gloo.  The per-rank compute is the CPU oracle (this is a test of the
data-parallel CONTRACT — sharding, global-batch gradient scaling, sum all-reduce of the flat bucket, identical SGD on every
rank — not of the CUDA kernels): two half-batch shards must reproduce the single-process step.
"""
import os
import socket

import numpy as np
import pytest

import oracle as O
from sliced_b200 import dp

DIMS = [24, 32, 16, 10]
BATCH = 64
LR = 0.1


def _problem():
    rng = np.random.default_rng(7)
    x = rng.uniform(0, 1, (BATCH, DIMS[0])).astype(np.float32)
    labels = rng.integers(0, 10, BATCH).astype(np.int32)
    y = np.zeros((BATCH, 10), np.float32)
    y[np.arange(BATCH), labels] = 1
    W = [rng.uniform(-0.1, 0.1, DIMS[i] * DIMS[i + 1]).astype(np.float32) for i in range(3)]
    B = [np.zeros(DIMS[i + 1], np.float32) for i in range(3)]
    return x, y, labels, W, B


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, labels, W, B = _problem()
    lo, hi = dp.shard_rows(BATCH, world, rank)
    losses = []
    for _ in range(3):
        loss, correct, dW, dB = O.mlp_step(0, DIMS, x[lo:hi].ravel().copy(), y[lo:hi].ravel().copy(), labels[lo:hi].copy(), W, B, LR,
                                           grad_rows=dp.grad_rows(BATCH), apply_sgd=False, want_grads=True)
        bucket = torch.from_numpy(np.concatenate([g.ravel() for pair in zip(dW, dB) for g in pair]))
        dp.exchange(bucket, world, lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM))
        flat = bucket.numpy()
        off = 0
        for l in range(3):
            for p in (W[l], B[l]):
                O.sgd_step(p, flat[off:off + p.size].copy(), LR)
                off += p.size
        m = torch.tensor([loss, float(correct)], dtype=torch.float64)
        dist.all_reduce(m)
        losses.append(m.numpy().copy())
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), *W, *B, losses=np.array(losses))
    dist.destroy_process_group()


def test_shard_rows():
    assert dp.shard_rows(65536, 8, 3) == (24576, 32768)
    assert dp.shard_rows(65536, 1, 0) == (0, 65536)
    with pytest.raises(ValueError):
        dp.shard_rows(10, 3, 0)
    with pytest.raises(ValueError):
        dp.shard_rows(8, 2, 2)
    covered = sorted(dp.shard_rows(64, 4, r) for r in range(4))
    assert covered[0][0] == 0 and covered[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_two_rank_step_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # single process, full batch
    x, y, labels, W, B = _problem()
    ref_losses = []
    for _ in range(3):
        loss, correct, _, _ = O.mlp_step(0, DIMS, x.ravel().copy(), y.ravel().copy(), labels, W, B, LR)
        ref_losses.append([loss, correct])
    keys = [k for k in r0.files if k != "losses"]
    for k, ref in zip(keys, W + B):
        assert np.array_equal(r0[k], r1[k]), "replicas must stay bit-identical across ranks"
        assert np.max(np.abs(r0[k] - ref)) <= 1e-5 * max(np.max(np.abs(ref)), 1e-3)
    assert np.allclose(r0["losses"], np.array(ref_losses), rtol=1e-5)
