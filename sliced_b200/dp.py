"""
Data-parallel plumbing of the nn.rs training step — host-side logic only, no arithmetic:

  * shard_rows: rank g of G owns rows [g*B/G, (g+1)*B/G) of x and y; weights are replicated
  * grad_rows:  cce_grad divides by `rows` (examples/nn.rs:151) — under sharding every rank must divide by the GLOBAL batch
                so that the sum-all-reduce reproduces the single-GPU gradient
  * exchange:   one sum all-reduce of the flat gradient bucket per step (NCCL on the GPU; the same contract is exercised
                with gloo on CPU tensors in tests/test_dp_gloo.py)
  * init_comm:  ship the 128-byte NCCL id from rank 0 to everyone through torch.distributed, then sl_comm_init_rank
"""
from __future__ import annotations

import ctypes as C


def shard_rows(global_batch: int, world: int, rank: int) -> tuple[int, int]:
    """[begin, end) rows of this rank; the batch must divide evenly (the reference has no ragged-batch handling)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def grad_rows(global_batch: int) -> int:
    return global_batch


def exchange(bucket, world: int, allreduce_sum):
    """In-place sum of the gradient bucket over the ranks; a world of one is a no-op."""
    if world > 1:
        allreduce_sum(bucket)
    return bucket


def init_comm(lib, ctx_handle, dist, rank: int, world: int, device="cuda"):
    """Create the library's NCCL communicator on every rank (id travels over the already-initialised torch.distributed group)."""
    import torch
    from . import capi
    idbuf = (C.c_char * 128)()
    if rank == 0:
        capi.check(None, lib.sl_comm_unique_id(idbuf))
    t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).to(device)
    dist.broadcast(t, 0)
    capi.check(ctx_handle, lib.sl_comm_init_rank(ctx_handle, world, rank, bytes(t.cpu().numpy().tobytes())))
