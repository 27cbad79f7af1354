#!/usr/bin/env python
"""Encrypted mean / covariance with the data points sharded across GPUs (BASELINE.json config 3:
Test_Statistics on d=4, N=10000, p=1019, g=3).

Follows Statistics.h:46-128 and Test_Statistics.cpp:196-244 for the circuit and the parameters:
  per rank     encrypt own blocks of 256 points and the block sizes; partial sums of the columns
               (coefficient form), of the block sizes, and of X_i*X_j (tensor form)
  exchange     all-gather of the partial sums, modular adds (one step, latency bound)
  replicated   mean_j = slot-sum(sum_b X_bj); cov_ij = N * slot-sum(KS(sum_b X_bi X_bj)) - mean_i mean_j;
               N^2; masking noise
Checks the decrypted values against ComputeNthMomentPT / ComputeCovariancePT (Statistics.h:173-208)
mod p.  One JSON line; phase names as in the reference driver.

  python apps/statistics_sharded.py --dim 4 --points 10000
"""
import argparse
import os

# load every kernel when the CUDA context is created (process start-up, before the drivers' clock)
# instead of lazily at first launch, which would land in whichever phase uses a kernel first
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "apps"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from fhesi_app import (Ct, Env, Slots, add_mask, decrypt_slot0, embed_batch, encrypt_batch,  # noqa: E402
                       keyswitch_and_sum_slots_batch,
                       rotation_exponents, sum_slots)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", dest="d", type=int, default=4)
    ap.add_argument("--points", dest="n", type=int, default=10000)
    ap.add_argument("--prime", dest="p", type=int, default=1019)
    ap.add_argument("--gen", dest="g", type=int, default=3)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--cpu-tensors", action="store_true", help="host tensors + gloo (emulator tests only)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    device = "cpu" if args.cpu_tensors else f"cuda:{local}"
    if not args.cpu_tensors:
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo" if args.cpu_tensors else "nccl")

    import pyfhesi
    from generate_random_data import generate
    from pyfhesi.hostkeys import keygen
    from pyfhesi.sharded import shard_bounds
    if args.lib is None:
        import build as fhesi_build
        args.lib = fhesi_build.build()
    p, g, d, N = args.p, args.g, args.d, args.n
    m = p - 1
    t_load0 = time.perf_counter()
    rows, _ = generate(d, N, args.seed)
    raw = np.asarray(rows, dtype=np.int64)               # LoadData's Matrix<ZZ>
    nslots = (p - 1) // 2 - 1
    block = 1 << (((p - 1) // 2).bit_length() - 1)      # Test_Statistics.cpp:193-198
    block = min(block, 1 << (nslots.bit_length() - 1))  # never more than the usable slots
    nblocks = (N + block - 1) // block
    xi = max(nblocks, d)
    logq = int(math.ceil((6.5 * math.log(nslots) + math.log(xi)) / math.log(2) + 36.1))  # :216-217

    slots = Slots(m, p, g, [(-1) ** i for i in range(m // 2)])
    rot_k = rotation_exponents(g, m, slots.usable)
    dev = pyfhesi.Context(m, logq, p, 3, xi, 0 if args.cpu_tensors else local, lib_path=args.lib)
    if not args.cpu_tensors:
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        dev.set_stream(stream.cuda_stream)
    env = Env(dev, device, staging_bytes=(nblocks + 8) * (d + 1) * block * 4)
    from pyfhesi.hostkeys import prepare as prepare_host_layer
    prepare_host_layer(args.lib)  # build-if-stale + dlopen of the C++ host layer: start-up, not key generation
    if not args.cpu_tensors:  # load torch's generator kernels now: process start-up, like the CUDA context
        from fhesi_app import encryption_randomness
        encryption_randomness(env, 1, np.random.default_rng(0))
    if world > 1:  # the communicator is process start-up, like the CUDA context: created before the clock
        warm = torch.zeros(8, dtype=torch.int32, device=device)
        dist.all_gather([torch.empty_like(warm) for _ in range(world)], warm)
        if not args.cpu_tensors:
            torch.cuda.synchronize()
    dev.sync()
    # the reference's clock starts at `Statistics stats(context)` = key generation
    # (Test_Statistics.cpp:112-173)
    t_start = time.perf_counter()
    if os.environ.get("FHESI_HOST_KEYGEN"):  # key-switch arithmetic on the host (C++ layer), then upload
        keys = keygen(dev, args.seed, g, rot_k=rot_k, lib_path=args.lib)
        t_keygen = time.perf_counter()
        ksw = dev.ksw_create(keys["ks_b"], keys["ks_A"], 3)
        rot_ksw = [dev.ksw_create(keys["rot_b"][i], keys["rot_A"][i], 2) for i in range(len(rot_k))]
        dpk, dsk = dev.key_create(keys["pk"]), dev.key_create(keys["sk"])
    else:  # draws on the host in the reference's stream order (flat arrays, no big-integer temporaries); every
        # matrix and the public key in one pass of kernels on the device (fhesi_keygen_batch)
        from pyfhesi.hostkeys import keydraws_flat, sk_words
        draws = keydraws_flat(dev, args.seed, g, rot_k=rot_k, lib_path=args.lib)
        t_keygen = time.perf_counter()
        ksws, dpk = dev.keygen_batch(draws["parts"], draws["src"], draws["sk"], draws["A"], draws["e"], with_pk=True)
        ksw, rot_ksw = ksws[0], ksws[1:]
        dsk = dev.key_create(sk_words(dev, draws["sk"]))
    dev.sync()
    t_setup = time.perf_counter()

    # ---- Batch + Encryption (Test_Statistics.cpp:35-63, Statistics.h:29-42)
    lo, hi = shard_bounds(nblocks, rank, world)
    nb, n = hi - lo, dev.n
    data = np.zeros((nblocks * block, d), dtype=np.int64)
    data[:N] = raw
    mine = (data[lo * block:hi * block] % p).reshape(max(nb, 0), block, d).transpose(0, 2, 1)
    d_msgs = embed_batch(env, slots, np.ascontiguousarray(mine).reshape(nb * d, block))
    sizes = np.array([[min(N, (lo + b + 1) * block) - (lo + b) * block] for b in range(nb)], dtype=np.int64)
    size_msgs = np.zeros((max(nb, 1), n), dtype=np.int32)       # Plaintext(context, n): the constant n
    size_msgs[:nb, 0] = (sizes[:, 0] % p) if nb else 0
    t_batch = time.perf_counter()
    nrng = np.random.default_rng(args.seed + 1000 + rank)
    cts = encrypt_batch(env, dpk, d_msgs, nb * d, nrng)
    ncts = encrypt_batch(env, dpk, torch.from_numpy(size_msgs).to(device), nb, nrng)
    dev.sync()
    t_enc = time.perf_counter()

    # ---- per-rank partial sums, then one exchange
    cw, tw = dev.ct_words(2), dev.tprod_words(3)
    col = lambda j: cts.view(max(nb, 1), d, cw)[:nb, j].contiguous()
    lin = torch.zeros((d + 1, cw), dtype=torch.int32, device=device)       # sums of columns, of block sizes
    pairs = [(i, j) for i in range(d) for j in range(i, d)]
    quad = torch.zeros((len(pairs), tw), dtype=torch.int32, device=device)  # sums of X_i X_j, tensor form
    if nb:
        for j in range(d):
            dev.ct_sum_dev(col(j), lin[j], 2, nb)
        dev.ct_sum_dev(ncts[:nb].contiguous(), lin[d], 2, nb)
        for idx, (i, j) in enumerate(pairs):
            dev.ct_tensor_dev(col(i), 2, col(j), 2, quad[idx], nb, accumulate=True)
    dev.sync()
    if world > 1:
        g_lin = torch.empty((world,) + tuple(lin.shape), dtype=torch.int32, device=device)
        g_quad = torch.empty((world,) + tuple(quad.shape), dtype=torch.int32, device=device)
        dist.all_gather([g_lin[w] for w in range(world)], lin)
        dist.all_gather([g_quad[w] for w in range(world)], quad)
        for k in range(d + 1):  # coefficient-form partial sums: Reduce(sum over ranks)
            dev.ct_sum_dev(g_lin[:, k].contiguous(), lin[k], 2, world)
        tot = torch.empty_like(quad)
        dev.tprod_reduce_gathered_dev(g_quad, world, 3 * len(pairs), tot)
        quad = tot
        dev.sync()
    t_data = time.perf_counter()

    # ---- replicated tail (Statistics.h:46-128)
    mean = [sum_slots(Ct(env, lin[j].clone(), 2), rot_k, rot_ksw) for j in range(d)]
    for ct in mean:
        add_mask(env, dpk, slots, ct, nrng)
    n_ct = Ct(env, lin[d].clone(), 2)
    mu = {}
    for (i, j) in pairs:
        t = mean[i].copy().mul(mean[j])
        t.keyswitch_(ksw)
        mu[(i, j)] = t.neg_()
    cov = {}
    summed = keyswitch_and_sum_slots_batch(env, quad, len(pairs), ksw, rot_k, rot_ksw)
    for idx, (i, j) in enumerate(pairs):
        c = summed[idx]
        c = c.mul(n_ct)
        c.keyswitch_(ksw)
        c.add_(mu[(i, j)])
        cov[(i, j)] = add_mask(env, dpk, slots, c, nrng)
    n2 = n_ct.copy().mul(n_ct)
    n2.keyswitch_(ksw)
    dev.sync()
    t_comp = time.perf_counter()

    got_mean = [decrypt_slot0(env, dsk, slots, c) for c in mean]
    got_n = decrypt_slot0(env, dsk, slots, n_ct)
    got_cov = [decrypt_slot0(env, dsk, slots, cov[pq]) for pq in pairs]
    got_n2 = decrypt_slot0(env, dsk, slots, n2)
    t_dec = time.perf_counter()

    # ---- plaintext check (Statistics.h:173-208), mod p
    X = np.asarray(rows, dtype=object)
    s1 = [int(sum(int(r[j]) for r in rows)) for j in range(d)]
    want_mean = [v % p for v in s1]
    want_cov = [(N * int(sum(int(r[i]) * int(r[j]) for r in rows)) - s1[i] * s1[j]) % p for (i, j) in pairs]
    # the masked means carry noise outside slot 0 only; the covariance uses the *masked* means exactly as
    # Statistics.h:86-100 does (noise lands in the other slots of mu_i * mu_j, never in slot 0)
    ok = (got_mean == want_mean and got_n == N % p and got_cov == want_cov and got_n2 == (N % p) ** 2 % p)
    if rank == 0:
        print(f"Setup time: {t_setup - t_start:.3f}\nBatch time: {t_batch - t_setup:.3f}\n"
              f"Encryption time: {t_enc - t_batch:.3f}\nComputation time: {t_comp - t_enc:.3f} "
              f"(partial sums + exchange {t_data - t_enc:.3f})\nDecryption time: {t_dec - t_comp:.3f}\n"
              f"Total time: {t_dec - t_start:.3f}")
        print(json.dumps({
            "metric": f"Test_Statistics N={N} d={d} wall time", "unit": "s", "higher_is_better": False,
            "value": t_dec - t_start, "n_gpus": world, "correct": bool(ok),
            "clock": "Test_Statistics.cpp:112-173 (key generation .. decryption)", "load_context_and_communicator_s": t_start - t_load0,
            "setup_split_s": {"host_draws_or_keygen": t_keygen - t_start, "device_generation_or_upload": t_setup - t_keygen},
            "phases_s": {"setup": t_setup - t_start, "batch": t_batch - t_setup, "encryption": t_enc - t_batch,
                         "partial_sums_and_exchange": t_data - t_enc, "replicated_tail": t_comp - t_data,
                         "decryption": t_dec - t_comp},
            "config": {"p": p, "g": g, "logQ": logq, "xi": xi, "blocks": nblocks, "block_size": block,
                       "chain": f"{dev.Lt}/{dev.Lk}", "blocks_per_rank": nb},
            "mean": got_mean, "N": got_n, "cov_upper": got_cov, "N2": got_n2,
            "expected": {"mean": want_mean, "N": N % p, "cov_upper": want_cov, "N2": (N % p) ** 2 % p}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit(f"rank {rank}: decrypted statistics differ from the plaintext computation")


if __name__ == "__main__":
    main()
