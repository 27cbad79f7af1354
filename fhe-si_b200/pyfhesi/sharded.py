"""Multi-GPU sharding of independent ciphertexts.

The only exchange step the workload has: every rank accumulates the tensor products of its
own shard of data blocks into ONE partial sum in tensor (tprod) form -- sums mod p_i are
associative and commutative, so any sharding is bit-identical to the serial order of
Matrix.cpp:80-97,149-173 -- then the partial sums are all-gathered (NCCL over NVLink on the
GPU box, gloo in CPU tests) and combined by a modular-add kernel
(fhesi_tprod_reduce_gathered_dev).  One process per GPU; torch.distributed is plumbing.

The mult+relin throughput sweep needs none of this: ranks own disjoint batches and never
communicate (bench.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(count: int, rank: int, world: int):
    """Contiguous near-equal split: 391 blocks over 8 ranks -> 49,49,49,49,49,49,49,48."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_tensor_sum(ctx, a: torch.Tensor, b: torch.Tensor, parts_a: int = 2, parts_b: int = 2,
                       group=None) -> torch.Tensor:
    """sum over ALL ranks' local pairs of a_i * b_i, in tprod form, replicated on every rank.

    a, b: this rank's shard, int32/uint32 tensors [count_local][parts][n][W] on the device
    the context lives on (CPU tensors when driven by the test emulator).  Returns an int32
    tensor [parts_a+parts_b-1][Lt][N]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    po = parts_a + parts_b - 1
    count_local = a.shape[0]
    local = torch.zeros((po, ctx.Lt, ctx.N), dtype=torch.int32, device=a.device)
    if count_local:
        ctx.ct_tensor_dev(a, parts_a, b, parts_b, local, count_local, accumulate=True)
    ctx.sync()
    if world == 1:
        return local
    gathered = torch.empty((world, po, ctx.Lt, ctx.N), dtype=torch.int32, device=a.device)
    dist.all_gather([gathered[w] for w in range(world)], local, group=group)  # NCCL / gloo
    out = torch.empty_like(local)
    ctx.tprod_reduce_gathered_dev(gathered, world, po, out)
    ctx.sync()
    return out
