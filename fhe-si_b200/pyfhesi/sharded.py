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
    if local.is_cuda:  # the fill ran on torch's current stream, the accumulation runs on the context's
        torch.cuda.current_stream(local.device).synchronize()
    if count_local:
        ctx.ct_tensor_dev(a, parts_a, b, parts_b, local, count_local, accumulate=True)
    ctx.sync()
    if world == 1:
        return local
    return allgather_add(ctx, local, po, group)


def allgather_add(ctx, local: torch.Tensor, parts: int, group=None) -> torch.Tensor:
    """The exchange step on its own: all-gather every rank's `local` partial sums (tprod form, any
    leading shape, `parts` * count polynomials of Lt x N words in all) and add them mod p_i.

    Stream discipline: the caller's kernels ran on the context's stream, the collective runs on
    torch's current stream, the combine kernel on the context's stream again -- which may all be
    different streams (a Context owns a non-blocking stream unless set_stream() was called).  Both
    hand-overs are made explicit here: ctx.sync() before the collective, and a synchronise of torch's
    current stream after it, so the combine never reads `gathered` early."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    ctx.sync()
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(gathered, local.contiguous(), group=group)  # one ncclAllGather
        torch.cuda.current_stream(local.device).synchronize()
    else:
        dist.all_gather([gathered[w] for w in range(world)], local, group=group)  # gloo (CPU tests)
    out = torch.empty_like(local)
    ctx.tprod_reduce_gathered_dev(gathered, world, parts, out)
    ctx.sync()
    return out
