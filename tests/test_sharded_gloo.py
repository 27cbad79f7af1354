"""world_size-2 gloo test of the N>1 path: shard the data-phase tensor
products across ranks, all-gather the partial tprod sums, combine with the modular-add kernel,
ScaleDown + key switch, and compare with the oracle's serial sum.  CPU: the kernels run in the
test emulator; on the GPU box the same code runs over NCCL (bench/test with --gpus 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, nblocks, results):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fhe-si_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fhesi_oracle as O
        from common import CONFIGS, Scenario, assert_ct_equal
        from pyfhesi.sharded import shard_bounds, sharded_tensor_sum
        # parameters sized from the GLOBAL block count, identical on all ranks
        sc = Scenario(*CONFIGS["cfg1"], seed=99, xi=nblocks, lib_path=emu_lib)
        _, cts = sc.fresh(2 * nblocks)  # same stream on every rank -> same global data set
        A, B = cts[:nblocks], cts[nblocks:]
        lo, hi = shard_bounds(nblocks, rank, world)
        d = sc.dev
        shape = (hi - lo, 2, d.n, d.W)
        ta = torch.from_numpy(sc.pack_cts(A[lo:hi]).view(np.int32)) if hi > lo else torch.zeros(shape, dtype=torch.int32)
        tb = torch.from_numpy(sc.pack_cts(B[lo:hi]).view(np.int32)) if hi > lo else torch.zeros(shape, dtype=torch.int32)
        total = sharded_tensor_sum(d, ta.contiguous(), tb.contiguous())
        c3 = torch.empty((3, d.n, d.W), dtype=torch.int32)
        d.scaledown_dev(total, 3, c3, 1)
        out = torch.empty((2, d.n, d.W), dtype=torch.int32)
        d.keyswitch_dev(sc.ksw, c3, out, 1)
        d.sync()
        acc = A[0].copy().mul(B[0])
        for i in range(1, nblocks):
            acc.add(A[i].copy().mul(B[i]))
        want = O.apply_key_switch(sc.ks, acc)
        assert_ct_equal(sc, out.numpy().view(np.uint32), want, f"rank {rank}")
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from pyfhesi.sharded import shard_bounds
    sizes = [shard_bounds(391, r, 8) for r in range(8)]
    assert [b - a for a, b in sizes] == [49] * 7 + [48]
    assert sizes[0][0] == 0 and sizes[-1][1] == 391
    assert all(sizes[i][1] == sizes[i + 1][0] for i in range(7))
    assert shard_bounds(1, 1, 2) == (1, 1)  # an empty shard is legal


@pytest.mark.parametrize("nblocks", [5, 1])
def test_sharded_sum_world2_gloo(emu_lib, nblocks):
    world = 2
    import socket
    with socket.socket() as sk:  # an unused rendezvous port
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, emu_lib, nblocks, results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
