#!/usr/bin/env python
"""Encrypted least-squares regression with the data points sharded across GPUs
(BASELINE.json config 4).

The reference has no multi-file / multi-process driver (Test_Regression.cpp takes one file and
sizes logQ from that file's N), so this is new code.  It follows Regression.h:102-149 and
Matrix.cpp:80-97,149-262 for the circuit:

  data phase   every rank encrypts its own blocks of 256 data points and accumulates its
               partial sums of X_i*y and X_i*X_j in tensor form (d + d(d+1)/2 = 14 sums, d = 4)
  exchange     one all-gather of the partial sums + modular add (pyfhesi.sharded)
  serial tail  key switch + slot sums by rotations, adjugate inverse with a key switch after
               every product level, adj * X^T y, masking noise -- replicated on every rank

Parameters come from the GLOBAL N: logQ by Test_Regression.cpp:107-108, xi = max(blocks, d).
Prints the reference's phase names and one JSON line; checks the decrypted theta*det and det
against the plaintext computation mod p.

  python apps/regression_sharded.py --dim 4 --points 100000                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... apps/regression_sharded.py --dim 4 --points 100000
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
for p in (os.path.join(ROOT, "fhe-si_b200"), os.path.join(ROOT, "scripts"), os.path.join(ROOT, "apps"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


from fhesi_app import (Ct, Env, Slots, _tensor_sum_keyswitch, adjugate_and_det_batched, embed_batch,  # noqa: E402
                       encrypt_batch, keyswitch_and_sum_slots_batch)


def determinant(M, rows, cols, reduce):
    """Matrix<T>::Determinant (Matrix.cpp:223-262): Laplace expansion along the first free row,
    products summed in tensor form, one `reduce` (key switch) per level."""
    row = rows[0]
    if len(rows) == 1:
        return M[row][cols[0]].copy()
    det, negative = None, False
    for col in cols:
        tmp = M[row][col].copy()
        if negative:
            tmp.neg_()
        negative = not negative
        minor = determinant(M, rows[1:], [c for c in cols if c != col], reduce)
        tmp = tmp.mul(minor)
        det = tmp if det is None else det.add_(tmp)
    reduce(det)
    return det


def plaintext_regression(rows, labels, p):
    """RegressPT (Regression.h:191-214): theta*det = adj(X^T X) X^T y and det, exact, mod p."""
    d = len(rows[0])
    A = [[sum(r[i] * r[j] for r in rows) for j in range(d)] for i in range(d)]
    b = [sum(r[i] * l for r, l in zip(rows, labels)) for i in range(d)]

    def det(M):
        if len(M) == 1:
            return M[0][0]
        return sum((-1) ** j * M[0][j] * det([r[:j] + r[j + 1:] for r in M[1:]]) for j in range(len(M)))
    if d == 1:
        return [b[0] % p], A[0][0] % p
    adj = [[(-1) ** (i + j) * det([r[:i] + r[i + 1:] for k, r in enumerate(A) if k != j]) for j in range(d)]
           for i in range(d)]
    theta = [sum(adj[i][k] * b[k] for k in range(d)) % p for i in range(d)]
    return theta, det(A) % p


def load_shard_files(prefix):
    """LoadData (Regression.h:14-41) over the files `generateRandomData.py name d N nFiles` writes
    (README:82-84): prefix_0.dat, prefix_1.dat, ... (or prefix.dat alone).  -> list of int64 arrays
    [rows_k][d + 1] (features then label), in file order."""
    paths, k = [], 0
    while os.path.exists(f"{prefix}_{k}.dat"):
        paths.append(f"{prefix}_{k}.dat")
        k += 1
    if not paths and os.path.exists(prefix + ".dat"):
        paths = [prefix + ".dat"]
    if not paths:
        raise SystemExit(f"no data files {prefix}_0.dat ... or {prefix}.dat")
    out = []
    for path in paths:
        with open(path) as f:
            d, n = (int(v) for v in f.readline().split())
            a = np.loadtxt(f, dtype=np.int64, ndmin=2) if n else np.zeros((0, d + 1), np.int64)
        assert a.shape == (n, d + 1), f"{path}: header says {n} x {d + 1}, file holds {a.shape}"
        out.append(a)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", dest="d", type=int, default=4)
    ap.add_argument("--points", dest="n", type=int, default=100000)
    ap.add_argument("--prime", dest="p", type=int, default=1019)
    ap.add_argument("--gen", dest="g", type=int, default=3)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--data", default=None, help="read NAME_0.dat, NAME_1.dat, ... (scripts/generate_random_data.py "
                    "NAME d N nFiles) instead of generating the data in-process; files are dealt to the ranks "
                    "round-robin and each file is cut into its own blocks, as one reference process per file would")
    ap.add_argument("--repeat", type=int, default=1, help="run the whole regression this many times in one process "
                    "(a server answering requests); every run is printed, a final line summarises them")
    ap.add_argument("--lib", default=None, help="C-ABI library (default: the in-tree CUDA build)")
    ap.add_argument("--cpu-tensors", action="store_true", help="host tensors + gloo (emulator tests only)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not args.cpu_tensors:
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo" if args.cpu_tensors else "nccl")
    runs = [run(args, rank, world, local) for _ in range(max(1, args.repeat))]
    res = min(runs, key=lambda r: r["value"])
    if args.repeat > 1 and rank == 0:
        print(json.dumps(dict(res, runs_s=[r["value"] for r in runs], runs_phases_s=[r["phases_s"] for r in runs])),
              flush=True)
    if not all(r["correct"] for r in runs):
        res = [r for r in runs if not r["correct"]][0]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not res["correct"]:
        raise SystemExit(f"rank {rank}: decrypted regression differs from the plaintext computation: "
                         f"{res['theta_det']} vs {res['expected']}")


def run(args, rank, world, local, quiet=False):
    """One encrypted regression; the process group (world > 1) exists already.  -> the result dict
    (rank 0 also prints it unless `quiet`)."""
    device = "cpu" if args.cpu_tensors else f"cuda:{local}"
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
    nslots = (p - 1) // 2 - 1
    block = 1 << (nslots.bit_length() - 1)              # Test_Regression.cpp:86-91
    if getattr(args, "data", None):
        files = load_shard_files(args.data)             # every rank reads the headers; LoadData is outside the clock
        d = files[0].shape[1] - 1
        N = sum(len(f) for f in files)
        raw = np.concatenate(files)
        # one reference process per file would cut each file into its own blocks (the last one ragged)
        file_blocks = [(len(f) + block - 1) // block for f in files]
        nblocks = sum(file_blocks)
    else:
        rows_, labels_ = generate(d, N, args.seed)      # every rank derives the same global data set
        raw = np.empty((N, d + 1), dtype=np.int64)      # LoadData's Matrix<ZZ> rawData + labels
        raw[:, :d] = np.asarray(rows_, dtype=np.int64)
        raw[:, d] = np.asarray(labels_, dtype=np.int64)
        files, file_blocks = [raw], [(N + block - 1) // block]
        nblocks = file_blocks[0]
    rows, labels = raw[:, :d].tolist(), raw[:, d].tolist()
    # LoadData's result in the width the device takes (values are small: |x| <= 100, |label| < 2^31 checked here)
    assert np.abs(raw).max(initial=0) < 2**31
    raw32 = np.ascontiguousarray(raw, dtype=np.int32)
    files32 = [np.ascontiguousarray(f, dtype=np.int32) for f in files]
    xi = max(nblocks, d)
    lgq = 4.5 * math.log(nslots) + max(1, d - 1) * (math.log(1280) + 2 * math.log(nslots) + math.log(xi))
    logq = int(math.ceil(lgq / math.log(2) + 24.7))     # Test_Regression.cpp:107-108

    # ---- LoadData + FHEcontext + SetUpSIContext: before the reference's clock starts
    # (Test_Regression.cpp:95-137); reported separately as "load_and_context"
    assert m % 2 == 0, "m = p - 1 must be 2 * (odd prime)"
    phi = [(-1) ** i for i in range(m // 2)]           # Phi_m(X) = sum (-1)^i X^i for m = 2p'
    slots = Slots(m, p, g, phi)
    rot_k, k, ns = [], g % m, slots.usable
    while ns > 1:                                       # Regression.h:70-81
        rot_k.append(k)
        ns >>= 1
        k = k * k % m
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
    import gc
    gc.collect()  # a previous run's context (cudaFree of its pool: device-wide synchronisations) must not be
    gc.disable()  # collected in the middle of this run's clock
    dev.sync()
    # ---- Setup = `Regression regress(context)` (Test_Regression.cpp:24-26, Regression.h:68-81): secret
    # and public key, s^2 and rotation key-switch matrices (C++ host layer), upload.  "Total time"
    # runs from here to the end of decryption, as in the reference driver (:24,:63).
    t_start = time.perf_counter()
    if os.environ.get("FHESI_HOST_KEYGEN"):  # key-switch arithmetic on the host (C++ layer), then upload
        keys = keygen(dev, args.seed, g, rot_k=rot_k, lib_path=args.lib)
        t_keygen = time.perf_counter()
        ksw = dev.ksw_create(keys["ks_b"], keys["ks_A"], 3)
        rot_ksw = [dev.ksw_create(keys["rot_b"][i], keys["rot_A"][i], 2) for i in range(len(rot_k))]
        dpk, dsk = dev.key_create(keys["pk"]), dev.key_create(keys["sk"])
    else:  # draws on the host in the reference's stream order (flat arrays, no big-integer temporaries); every
        # matrix and the public key in one pass of kernels on the device (fhesi_keygen_batch).  The draws are
        # pure host work (a C call that releases the interpreter lock): they run on a thread of their own while
        # this thread batches the data below -- BatchData needs no key -- and are joined before encryption.
        import threading
        from pyfhesi.hostkeys import keydraws_flat, sk_words
        box = {}

        def draw():
            box["draws"] = keydraws_flat(dev, args.seed, g, rot_k=rot_k, lib_path=args.lib)
            box["t"] = time.perf_counter()
        draw_thread = threading.Thread(target=draw)
        draw_thread.start()
        t_keygen = None
    if t_keygen is not None:
        dev.sync()
    t_setup = time.perf_counter()

    # ---- Batch + Encryption of this rank's blocks (BatchData, Regression.h:43-66; AddData :83-95)
    n = dev.n
    # all of this rank's plaintexts at once: [nb][d+1][block] slot values -> PlaintextSpace::EmbedInSlots
    # on the device (fhesi_embed_slots_dev, exact integer arithmetic)
    # (one strided copy per file, straight into the final [block][column][slot] layout in 32-bit words: the values
    # are small integers -- the first version padded, concatenated and transposed 64-bit copies, 6 ms of a 17 ms clock)
    def pack(dst, rows):                                # rows [L][d+1] -> dst [ceil(L / block)][d+1][block], zero padded
        full = len(rows) // block
        if full:
            dst[:full] = rows[:full * block].reshape(full, block, d + 1).transpose(0, 2, 1)
        if len(rows) > full * block:
            dst[full, :, :len(rows) - full * block] = rows[full * block:].T
    if len(files) >= world and len(files) > 1:          # whole files per rank, round-robin
        own = list(range(rank, len(files), world))
        nb = sum(file_blocks[k] for k in own)
        vals3 = env.staging_array((nb, d + 1, block))
        b0 = 0
        for k in own:
            pack(vals3[b0:b0 + file_blocks[k]], files32[k])
            b0 += file_blocks[k]
    else:                                               # fewer files than ranks: split the global block list
        lo, hi = shard_bounds(nblocks, rank, world)
        nb = max(hi - lo, 0)
        vals3 = env.staging_array((nb, d + 1, block))
        if nb:
            pack(vals3, raw32[lo * block:min(hi * block, N)])
    t_b0 = time.perf_counter()
    vals = vals3.reshape(nb * (d + 1), block)
    t_b1 = time.perf_counter()
    d_msgs = embed_batch(env, slots, vals)
    t_b2 = time.perf_counter()
    dev.sync()
    t_batch = time.perf_counter()
    if os.environ.get("FHESI_APP_TIMING"):
        print(f"batch: blocks {t_b0 - t_setup:.4f}  transpose {t_b1 - t_b0:.4f}  embed enqueue {t_b2 - t_b1:.4f}  "
              f"sync {t_batch - t_b2:.4f}", file=sys.stderr)
    if t_keygen is None:  # the keys: join the draws, then the device half (1 ms)
        draw_thread.join()
        draws = box["draws"]
        t_keygen = box["t"]
        ksws, dpk = dev.keygen_batch(draws["parts"], draws["src"], draws["sk"], draws["A"], draws["e"], with_pk=True)
        ksw, rot_ksw = ksws[0], ksws[1:]
        dsk = dev.key_create(sk_words(dev, draws["sk"]))
        dev.sync()
        t_keys = time.perf_counter()
        # report the phases as wall-clock segments: batch ran first (with the draws behind it), "setup" is what
        # key generation added after it
        t_setup, t_batch = t_start + (t_keys - t_batch), t_keys
    nrng = np.random.default_rng(args.seed + 1000 + rank)
    cnt = nb * (d + 1)
    cts = encrypt_batch(env, dpk, d_msgs, cnt, nrng)
    dev.sync()
    t_enc = time.perf_counter()

    # ---- data phase: partial sums in tensor form (Matrix.cpp:80-97, 149-173), then one exchange
    cw = dev.ct_words(2)
    col = lambda j: cts.view(max(nb, 1), d + 1, cw)[:nb, j].contiguous()
    pairs = [(i, d) for i in range(d)] + [(i, j) for i in range(d) for j in range(i, d)]
    po_words = dev.tprod_words(3)
    partial = torch.zeros((len(pairs), po_words), dtype=torch.int32, device=device)
    for idx, (i, j) in enumerate(pairs):
        if nb:
            dev.ct_tensor_dev(col(i), 2, col(j), 2, partial[idx], nb, accumulate=True)
    dev.sync()
    from pyfhesi.sharded import allgather_add
    total = allgather_add(dev, partial, 3 * len(pairs))  # the one exchange step (world == 1: a no-op)
    t_data = time.perf_counter()

    # ---- serial tail, replicated (Regression.h:106-148)
    def reduce(ct):
        ct.keyswitch_(ksw)

    sums = keyswitch_and_sum_slots_batch(env, total, len(pairs), ksw, rot_k, rot_ksw)
    S = torch.stack([c.buf for c in sums])                # [14][cw]: X^T y (d of them), then the upper triangle of X^T X
    xty = S[:d]
    tri = {}
    it = iter(range(d, len(sums)))
    for i in range(d):
        for j in range(i, d):
            tri[(i, j)] = tri[(j, i)] = next(it)
    E = S[torch.tensor([tri[(i, j)] for i in range(d) for j in range(d)], device=device)]  # X^T X, row-major
    if d == 1:
        det_b, theta_b = E[0:1].clone(), xty[0:1].clone()
    else:
        # Matrix::Invert (adjugate, Matrix.cpp:181-213) level by level, then dataCopy *= last; MapAll(KS)
        adj, det_b = adjugate_and_det_batched(env, E, d, ksw)
        A = adj.view(d, d, -1)
        theta_b = _tensor_sum_keyswitch(env, ksw, [(A[:, kx], xty[kx:kx + 1].expand(d, -1)) for kx in range(d)])
    res_b = torch.cat([theta_b, det_b])                   # theta_0 .. theta_{d-1}, det
    # masking noise in every slot but the first (Regression.h:180-189), all d + 1 ciphertexts at once
    nv = np.zeros((d + 1, slots.total), dtype=np.int64)
    nv[:, 1:] = nrng.integers(0, p, size=(d + 1, slots.total - 1))
    noise = encrypt_batch(env, dpk, embed_batch(env, slots, nv), d + 1, nrng)
    dev.ct_add_dev(res_b, noise, 2, d + 1)
    dev.sync()
    t_reg = time.perf_counter()

    # ---- Decryption: one batched call, one read-back
    mbuf = torch.empty((d + 1, n), dtype=torch.int32, device=device)
    dev.decrypt_dev(dsk, res_b, 2, mbuf, d + 1)
    dev.sync()
    hm = mbuf.cpu().numpy().view(np.uint32)
    out = [slots.decode0(hm[i]) for i in range(d + 1)]
    t_dec = time.perf_counter()
    gc.enable()

    want_theta, want_det = plaintext_regression(rows, labels, p)
    ok = out[:-1] == want_theta and out[-1] == want_det
    total_s = t_dec - t_start
    if world > 1:  # wall time of the job = the slowest rank
        tt = torch.tensor([total_s], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_s = float(tt.item())
    res = {
        "metric": f"Test_Regression N={N} d={d} wall time", "unit": "s", "higher_is_better": False,
        "value": total_s, "n_gpus": world, "correct": bool(ok),
        "clock": "Test_Regression.cpp:24-63 (key generation .. decryption), max over ranks",
        "load_context_and_communicator_s": t_start - t_load0,
        "setup_split_s": {"host_draws_or_keygen": t_keygen - t_start,
                          "note": "the host draws run on their own thread behind the batch phase; phases_s.setup is the "
                                  "part of key generation that was not hidden (join + device generation)"},
        "phases_s": {"setup": t_setup - t_start, "batch": t_batch - t_setup, "encryption": t_enc - t_batch,
                     "data_phase_and_exchange": t_data - t_enc, "serial_tail": t_reg - t_data,
                     "decryption": t_dec - t_reg},
        "config": {"p": p, "g": g, "logQ": logq, "xi": xi, "blocks": nblocks, "block_size": block,
                   "chain": f"{dev.Lt}/{dev.Lk}", "blocks_per_rank": nb,
                   "input": (f"{len(files)} shard files {os.path.basename(args.data)}_k.dat" if getattr(args, "data", None)
                             else "generated in-process")},
        "theta_det": out, "expected": want_theta + [want_det]}
    # deterministic tear-down, outside the clock: key images and the context's buffer pool
    for k in [ksw] + list(rot_ksw):
        dev.lib.fhesi_ksw_destroy(k)
    dev.lib.fhesi_key_destroy(dpk)
    dev.lib.fhesi_key_destroy(dsk)
    del env, cts, partial, total, sums, res_b
    dev.close()
    if rank == 0 and not quiet:
        print(f"Setup time: {t_setup - t_start:.3f}\nBatch time: {t_batch - t_setup:.3f}\n"
              f"Encryption time: {t_enc - t_batch:.3f}\nRegression time: {t_reg - t_enc:.3f} "
              f"(data phase + exchange {t_data - t_enc:.3f})\nDecryption time: {t_dec - t_reg:.3f}\n"
              f"Total time: {t_dec - t_start:.3f}")
        print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    main()
