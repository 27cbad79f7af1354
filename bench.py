#!/usr/bin/env python
"""bench.py -- batched ciphertext mult+relinearise throughput (BASELINE.json metric).

A step is one pass of the hot path (c = a; c *= b; ks.ApplyKeySwitch(c), Test_AddMul.cpp:59-66)
over one batch of B independent fresh ciphertext pairs per GPU at the cfg2 parameters
(logQ=256, p=1019, g=3 -- g=2 is not in Z_1018^*).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...             # CPU arm: the reference algorithm

Under torchrun (N>1) every rank drives its own GPU over its own disjoint batch (no data-path
collective: the units are independent); time is the max over ranks.

The same JSON line carries BASELINE.json's second metric under "regression": Test_Regression on
d=4, N=100000 split into 8 files (config 4), wall time on the reference's clock
(Test_Regression.cpp:24-63), through apps/regression_sharded.py on the same N GPUs -- the data
blocks sharded over the ranks, one NCCL all-gather + modular add, replicated tail -- with the
reference's own Regression.h (oracle/_ref/ref_regression, CPU) timed beside it.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fhe-si_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

# every kernel of the library is loaded when the CUDA context is created, not on its first launch inside a timed
# region (the regression leg touches some forty kernels the throughput legs never launch)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np  # noqa: E402

SEED = 20240611
CFG = {"logQ": 256, "p": 1019, "g": 3}
METRIC = "ciphertext mult+relin/sec at logQ=256"

# canonical algorithmic work per op (SURVEY.md §8d / BASELINE.md §5), 64-bit modmul-equivalents
CANON_MODMUL = {128: 1_759_670, 256: 5_514_040, 512: 17_824_284, 80: 14_322, 100: 1_248_160, 176: 3_022_054}


def algorithmic_bytes_per_op(n, logq):
    return 6 * n * (logq // 8)  # read 2 ct x 2 parts, write 2 parts (SURVEY.md §8d B_min)


# ---------------------------------------------------------------------------------------
# CPU arm: oracle/ref_restate.c, the reference algorithm restated in C (kind = "port")
# ---------------------------------------------------------------------------------------
def _cpu_setup(logq, p, g, faithful=False):
    import fhesi_oracle as O
    import ref_port
    octx = O.Context(p - 1, logq, p, g).setup_si()
    rng = O.Rng(SEED)
    sk = O.SecKey.generate(octx, rng)
    pk = O.PubKey.generate(sk, rng)
    ks = O.KeySwitch.init_s2(sk, rng)
    port = ref_port.RefPort(octx, faithful=faithful)
    port.set_key_switch(ks)
    pack = lambda ct: np.stack([O.pack_poly_words(x, logq) for x in ct.parts])
    msgs = [[rng.random_bnd(p) for _ in range(octx.phim)] for _ in range(2)]
    a, b = (pack(O.encrypt_rng(pk, m, rng)) for m in msgs)
    out = port.mult_relin(a, b)  # warm-up op: cached Rb tables exist, as in the reference
    return port, a, b, out


_WORKER = {}


def _cpu_worker_init(logq, p, g, faithful):
    """Once per worker process: context, keys, operands, one warm-up op (cached tables)."""
    _WORKER["state"] = _cpu_setup(logq, p, g, faithful)


def _cpu_worker(nops):
    port, a, b, _ = _WORKER["state"]
    t = time.perf_counter()
    for _ in range(nops):
        port.mult_relin(a, b)
    return time.perf_counter() - t


def cpu_baseline_single_core(logq, p, g, budget_s=12.0):
    """1 core, table-bug-fixed and faithful variants."""
    port, a, b, _ = _cpu_setup(logq, p, g, False)
    t = time.perf_counter()
    port.mult_relin(a, b)
    one = time.perf_counter() - t
    nops = max(2, min(64, int(budget_s / max(one, 1e-6))))
    t = time.perf_counter()
    for _ in range(nops):
        port.mult_relin(a, b)
    fixed = nops / (time.perf_counter() - t)
    port.lib.ref_set_faithful(port.h, 1)
    nf = max(1, nops // 6)
    t = time.perf_counter()
    for _ in range(nf):
        port.mult_relin(a, b)
    faithful = nf / (time.perf_counter() - t)
    out = {"value": fixed, "unit": "ops/s", "cores": 1, "kind": "port",
           "sample": f"{nops} mult+relin ops at logQ={logq} p={p} (oracle/ref_restate.c: m-point Bluestein "
                     f"N=2^{(2 * (p - 1) - 1).bit_length()}, incremental bigint CRT), tables cached",
           "faithful_table_bug_ops_per_s": faithful, "faithful_sample_ops": nf}
    out.update(reference_sources_figure(logq, p, g))
    return out


def reference_sources_figure(logq, p, g, ops=3):
    """The reference's OWN sources (oracle/_ref, built against the NTL stand-in by
    oracle/build_ref.py) on the same op, when the prebuilt binary travelled with the snapshot.
    Reported beside the port, not instead of it: the stand-in's big integers are slower than
    NTL/GMP, so the port is the faster -- fairer -- CPU figure."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    if not os.path.exists(exe):
        return {}
    try:
        r = subprocess.run([exe, str(logq), str(p), str(g), str(ops), str(SEED)], capture_output=True, text=True,
                           timeout=180)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return {"reference_sources_on_ntl_standin": {"ops_per_s": j["ops_per_s"], "ops": j["ops"], "cores": 1,
                                                     "decrypt_ok": j["decrypt_ok"],
                                                     "what": "oracle/_ref/ref_bench: the reference's DoubleCRT/"
                                                             "Bluestein/key-switch code, NTL replaced by "
                                                             "oracle/ntl_compat"}}
    except Exception as e:  # a missing or broken side figure must not break the bench line
        return {"reference_sources_on_ntl_standin": {"error": str(e)[:200]}}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    logq, p, g = CFG["logQ"], CFG["p"], CFG["g"]
    # size one step to ~2 s of wall time
    port, a, b, _ = _cpu_setup(logq, p, g)
    t = time.perf_counter()
    port.mult_relin(a, b)
    one = time.perf_counter() - t
    per_worker = max(1, int(2.0 / one))
    ctxm = mp.get_context("fork")
    times = []
    with ctxm.Pool(cores, initializer=_cpu_worker_init, initargs=(logq, p, g, False)) as pool:
        pool.map(_cpu_worker, [1] * cores, chunksize=1)  # every worker is set up before the clock starts
        for s in range(args.warmup + args.steps):
            t = time.perf_counter()
            pool.map(_cpu_worker, [per_worker] * cores, chunksize=1)
            dt = time.perf_counter() - t
            if s >= args.warmup:
                times.append(dt)
    ops_per_step = per_worker * cores
    ms = 1e3 * sum(times) / len(times)
    value = ops_per_step / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": bench_config(args.batch, args.gpus),
        "sample": {"ops_per_step": ops_per_step,
                   "note": "one process per core, each set up (context, key matrix, cached tables) before "
                           "the timed steps; a step is ops only"},
        "cpu_baseline": dict({"value": value, "unit": "ops/s", "cores": cores, "kind": "port",
                              "sample": f"{ops_per_step} ops/step over {cores} processes, oracle/ref_restate.c: the "
                                        "reference's algorithm restated in C.  The reference's own sources also run "
                                        "here (oracle/_ref, NTL replaced by a stand-in) but several times slower per "
                                        "core than this port, so the port is the arm that is timed"},
                             **reference_sources_figure(logq, p, g)),
        "e2e": {"value": value, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_regression:
        ref = regression_reference(full_tail=True)
        line["regression"] = dict({"metric": REG_METRIC, "value": ref.get("value"), "unit": "s",
                                   "higher_is_better": False, "n_gpus": args.gpus,
                                   "config": {"d": REG["d"], "N": REG["n"], "p": REG["p"], "g": REG["g"],
                                              "logQ": REG["logQ"], "xi": REG["xi"], "blocks": 392}},
                                  cpu_baseline=ref)
    print(json.dumps(line), flush=True)


def bench_config(batch, world):
    """The `config` object, identical in both arms (the reference arm times a bounded sample of the same
    workload; what the sample was is in its cpu_baseline.sample)."""
    n, W = (CFG["p"] - 1) // 2 - 1, (CFG["logQ"] + 31) // 32
    return {"workload": workload_name(), "batch_per_gpu": batch, "global_batch": batch * world,
            "parallelism": f"independent ciphertext shards x{world}, no data-path collective",
            "l2": f"inputs {2 * batch * 2 * n * W * 4 / 2**20:.0f} MiB per step > 126 MB L2",
            "seed": SEED}


def workload_name():
    return (f"{'cfg2' if CFG['logQ'] == 256 else 'cfg5'}: TestAddMul logQ={CFG['logQ']} p={CFG['p']} g={CFG['g']} "
            f"(m={CFG['p'] - 1}, phi(m)={(CFG['p'] - 1) // 2 - 1}, D={(CFG['logQ'] + 23) // 24}), "
            "batched c=a; c*=b; ApplyKeySwitch(c) on fresh encryptions")


# ---------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_sm = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host buffer exists, so
    the end-to-end path's staging memory is local to the GPU's PCIe root (8 ranks otherwise share
    whichever node the launcher started them on).  Returns the node, or None when the topology is
    not exposed; never fatal."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ---------------------------------------------------------------------------------------
# BASELINE.json metric 2: Test_Regression, d=4, N=100000 in 8 files, p=1019, g=3 (config 4)
# ---------------------------------------------------------------------------------------
REG = {"d": 4, "n": 100000, "p": 1019, "g": 3, "files": 8, "seed": 12345, "logQ": 176, "xi": 391}
REG_METRIC = "Test_Regression N=1e5 wall time"


def _regression_files(tmpdir):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import generate_random_data
    prefix = os.path.join(tmpdir, "reg4")
    generate_random_data.main(["generate_random_data.py", prefix, str(REG["d"]), str(REG["n"]), str(REG["files"]),
                               "--seed", str(REG["seed"])])
    return prefix


def run_regression_ours(args, rank, local_rank, world, lib_path, runs=11):
    """apps/regression_sharded.py on the N GPUs of this job, reading the 8 shard files; the clock is the
    reference driver's (key generation .. decryption, Test_Regression.cpp:24-63), max over ranks.  The first run
    pays one-off costs (first allocation of every buffer, first use of every kernel at these sizes); all runs are
    reported, `value` is the MEDIAN of the runs after the first (a shared box shows occasional stalls of tens of
    milliseconds in host-side calls; the median is robust to them without hiding them)."""
    import tempfile
    import types
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "apps"))
    import regression_sharded
    tmp = None
    if rank == 0:
        tmp = tempfile.mkdtemp(prefix="fhesi_reg_")
        _regression_files(tmp)
    if world > 1:
        box = [tmp]
        dist.broadcast_object_list(box, src=0)
        tmp = box[0]
    a = types.SimpleNamespace(d=REG["d"], n=REG["n"], p=REG["p"], g=REG["g"], seed=REG["seed"],
                              data=os.path.join(tmp, "reg4"), lib=lib_path, cpu_tensors=False)
    res = [regression_sharded.run(a, rank, world, local_rank, quiet=True) for _ in range(runs)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    warm = sorted(res[1:] or res, key=lambda r: r["value"])
    best = warm[(len(warm) - 1) // 2]  # the median run (its phases are the ones reported)
    if not all(r["correct"] for r in res):
        raise SystemExit("bench: the encrypted regression does not decrypt to RegressPT mod p: %r" % (res[0],))
    return {"metric": REG_METRIC, "value": best["value"], "unit": "s", "higher_is_better": False, "n_gpus": world,
            "runs_s": [r["value"] for r in res], "best_s": warm[0]["value"], "first_run_s": res[0]["value"],
            "runs_phases_ms": [{k: round(v * 1e3, 2) for k, v in r["phases_s"].items()} for r in res],
            "value_is": "median of the runs after the first", "clock": best["clock"], "phases_s": best["phases_s"],
            "setup_split_s": best["setup_split_s"], "config": dict(best["config"], d=REG["d"], N=REG["n"]),
            "theta_det": best["theta_det"], "expected": best["expected"], "correct": True,
            "scaling": "strong (the 392 data blocks are dealt to the ranks; keys and the 4x4 tail are replicated)"}


def regression_reference(full_tail):
    """The reference's own Regression.h / Matrix.cpp / FHE-SI code (oracle/_ref/ref_regression: its sources
    compiled against the NTL stand-in, one thread -- the reference is single-threaded) at config 4's
    parameters (logQ=176, xi=391).  The N-independent phases (set-up, and with full_tail the key-switch /
    rotation / adjugate tail and decryption) run whole; the N-proportional ones (batch, encryption,
    data-phase products) are timed on the first block of file 0 and scaled to the 392 blocks."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_regression")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_regression was not built (no reference tree at build time)"}
    with tempfile.TemporaryDirectory(prefix="fhesi_regref_") as tmp:
        prefix = _regression_files(tmp)
        cmd = [exe, prefix + "_0.dat", str(REG["p"]), str(REG["g"]), str(REG["logQ"]), str(REG["xi"]), "1",
               str(REG["seed"]), "0" if full_tail else "1"]
        t = time.perf_counter()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
            j = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            return {"unavailable": "ref_regression failed: " + str(e)[:200]}
        wall = time.perf_counter() - t
    blocks = 392
    per_block = j["batch_s"] + j["encryption_s"] + j["data_phase_s"]
    out = {"kind": "reference", "cores": 1, "unit": "s",
           "what": "oracle/_ref/ref_regression: the reference's Regression.h, Matrix.cpp, FHE-SI.cpp, DoubleCRT.cpp ... "
                   "unmodified, NTL replaced by oracle/ntl_compat (slower than NTL/GMP; labelled, not corrected)",
           "sample": "set-up whole%s; batch + encryption + data-phase products on 1 block of 256 points, x %d blocks"
                     % (", tail and decryption whole" if full_tail else "", blocks),
           "sample_wall_s": wall, "setup_s": j["setup_s"], "per_block_s": per_block,
           "n_proportional_s_extrapolated": per_block * blocks}
    if full_tail:
        tail = j["regression_s"] - j["data_phase_s"]
        out.update({"tail_s": tail, "decryption_s": j["decryption_s"], "decrypt_ok": j["correct"],
                    "value": j["setup_s"] + per_block * blocks + tail + j["decryption_s"]})
    else:
        out.update({"tail_s": None, "value": j["setup_s"] + per_block * blocks,
                    "note": "lower bound: the N-independent tail (~250 key switches, 160 rotations, the 4x4 adjugate) "
                            "is not sampled in this bounded leg; `bench.py --impl reference` runs it whole"})
    return out


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = None if os.environ.get("FHESI_NO_NUMA_BIND") else bind_to_gpu_numa_node(local_rank)
    if world > 1:
        # NCCL's own log (NCCL_DEBUG=INFO: communicator, rank count, transports) is left at the level the
        # launcher asked for; it only moves off stdout -- which carries the one JSON line -- to stderr
        # NCCL honours NCCL_DEBUG_FILE only above the VERSION level, and at VERSION (the level this pool's boxes
        # run at when nothing is set) its "NCCL version" banner goes to stdout ahead of the JSON line
        # (profiles/r02_bench_8gpu_logq*.json).  An unset / VERSION level is therefore RAISED to WARN -- never lowered.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import build as fhesi_build
    import pyfhesi
    from pyfhesi.hostkeys import keygen
    lib_path = fhesi_build.build()
    logq, p, g = CFG["logQ"], CFG["p"], CFG["g"]

    # ---- BASELINE.json's second metric: Test_Regression d=4 N=100000 (config 4).  Its clock is 17 ms of mostly
    # host-side work and is sensitive to what else the box is doing.  Measured on fresh boxes: at one GPU the runs
    # repeat to within a millisecond when the leg comes FIRST and scatter between 17 and 70 ms after the throughput
    # legs (profiles/r02h_bench_1gpu_fresh_box.json); under torchrun it is the other way round -- while the ranks'
    # start-up is still paging the image in, the first leg scatters up to 0.4 s, after the throughput legs it repeats
    # (profiles/r02g_bench_2gpu.json, r02f_bench_8gpu.json).  Every run and its phases are in the JSON either way.
    regression = None
    if not args.no_regression and world == 1:
        regression = run_regression_ours(args, rank, local_rank, world, lib_path)

    dev = pyfhesi.Context(p - 1, logq, p, 3, 1, local_rank, lib_path=lib_path)
    # one explicit (non-default) stream for the library's kernels AND the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    dev.set_stream(stream.cuda_stream)
    # keys: the C++ host layer's FHESISecKey / FHESIPubKey / KeySwitchSI (set-up, not timed)
    keys = keygen(dev, SEED, g, lib_path=lib_path)
    ksw = dev.ksw_create(keys["ks_b"], keys["ks_A"], 3)
    dpk = dev.key_create(keys["pk"])
    dsk = dev.key_create(keys["sk"])
    n, W = dev.n, dev.W
    B = args.batch

    # synthetic operands: 2B fresh encryptions per rank, made on the device from explicit randomness
    nrng = np.random.default_rng(SEED + rank)
    msgs = nrng.integers(0, p, size=(2 * B, n), dtype=np.uint32)
    rs = nrng.integers(0, 2, size=(2 * B, n), dtype=np.uint8)
    es = np.rint(nrng.normal(0.0, 3.2, size=(2 * B, 2, n))).astype(np.int32)
    ct_words = 2 * n * W
    d_ct = torch.empty((2 * B, ct_words), dtype=torch.int32, device="cuda")
    d_out = torch.empty((B, ct_words), dtype=torch.int32, device="cuda")
    t_msgs, t_rs, t_es = (torch.from_numpy(x).cuda() for x in (msgs.view(np.int32), rs, es))
    dev.encrypt_dev(dpk, t_msgs, t_rs, t_es, d_ct, 2 * B)
    torch.cuda.synchronize()
    d_a, d_b = d_ct[:B], d_ct[B:]

    def step():
        dev.mult_relin_dev(ksw, d_a, d_b, d_out, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness guard inside the bench (no oracle on this arm): the Test_AddMul.cpp:84-86
    # identity -- mult+relin of two fresh encryptions must decrypt to the plaintext product
    step()
    d_dec = torch.empty((B, n), dtype=torch.int32, device="cuda")
    dev.decrypt_dev(dsk, d_out, 2, d_dec, B)
    torch.cuda.synchronize()
    h_ct = d_ct.cpu().numpy().view(np.uint32).reshape(2 * B, 2, n, W)
    h_out = d_out.cpu().numpy().view(np.uint32).reshape(B, 2, n, W)
    dec = d_dec.cpu().numpy().view(np.uint32)
    h = (p - 1) // 2
    for i in (0, 1, B // 2, B - 1):
        u = np.convolve(msgs[i].astype(np.int64), msgs[B + i].astype(np.int64))  # < 2^40, exact
        v = np.zeros(h, dtype=np.int64)
        v[:h] += u[:h]
        v[:len(u) - h] -= u[h:]                      # X^h = -1
        sign = np.where(np.arange(n) % 2 == 0, 1, -1)
        want = (v[:n] - sign * v[n]) % p             # Phi_m = sum (-1)^i X^i
        assert np.array_equal(dec[i].astype(np.int64), want), "bench: mult+relin does not decrypt to the product"

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    crt_fallbacks = dev.crt_fallbacks()  # since the context was created: warm-up + timed steps
    launches = dev.launches()
    prof = dev.profile_report()
    dev.profile_enable(False)
    sampler.stop_flag = True
    sampler.join()

    # ---- e2e: pinned host buffers -> H2D -> hot path -> D2H, all inside the timed region
    # (operands in write-combined page-locked memory, fhesi_host_alloc: the CPU only writes them; the result in
    # ordinary page-locked memory, the CPU reads it -- profiles/r02c_pcie_probe_8gpu.txt)
    wc = not os.environ.get("FHESI_BENCH_NO_WC")
    np_a = dev.host_alloc((B, ct_words), np.int32, write_combined=wc)
    np_b = dev.host_alloc((B, ct_words), np.int32, write_combined=wc)
    np_a[...] = h_ct[:B].view(np.int32).reshape(B, ct_words)
    np_b[...] = h_ct[B:].view(np.int32).reshape(B, ct_words)
    h_a, h_b = torch.from_numpy(np_a), torch.from_numpy(np_b)
    h_o = torch.empty((B, ct_words), dtype=torch.int32).pin_memory()

    h_o2 = torch.empty((B, ct_words), dtype=torch.int32).pin_memory()

    def e2e_step():  # the blocking call: returns when the step's result is in host memory
        dev._ck(dev.lib.fhesi_mult_relin_host(dev.h, ksw, h_a.data_ptr(), h_b.data_ptr(), h_o.data_ptr(), B))

    def e2e_step_async(i):  # a server's loop: step i + 1 is enqueued while step i drains; results alternate buffers
        dev._ck(dev.lib.fhesi_mult_relin_host_async(dev.h, ksw, h_a.data_ptr(), h_b.data_ptr(),
                                                    (h_o, h_o2)[i & 1].data_ptr(), B))

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    e2e_steps = max(2, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_blocking_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps  # host-blocking call: wall clock
    assert np.array_equal(h_o.numpy().view(np.uint32).reshape(B, 2, n, W), h_out), "e2e result differs"
    h_o.zero_()
    for i in range(2):
        e2e_step_async(i)
    dev.sync_all()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step_async(i)
    dev.sync_all()  # every step's result is in host memory
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    for buf in (h_o, h_o2):
        assert np.array_equal(buf.numpy().view(np.uint32).reshape(B, 2, n, W), h_out), "e2e (async) result differs"

    # ---- the host's share of the end-to-end figure, measured: the same bytes per step as the e2e call
    # moves, copied concurrently in both directions on every rank at once (no kernels).  N ranks share one
    # host; this is the ceiling the host path sets on e2e at this N.
    cs1, cs2 = torch.cuda.Stream(), torch.cuda.Stream()
    d_pa, d_po = torch.empty_like(d_ct), torch.empty_like(d_out)

    def copy_step():
        with torch.cuda.stream(cs1):
            d_pa[:B].copy_(h_a, non_blocking=True)
            d_pa[B:].copy_(h_b, non_blocking=True)
        with torch.cuda.stream(cs2):
            h_o.copy_(d_po, non_blocking=True)

    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copy_step()
    barrier()
    pcie_ms = (time.perf_counter() - t0) * 1e3 / 3
    del d_pa, d_po

    # ---- the data-phase exchange on its own (N > 1): all-gather of every rank's 14 partial sums in tensor
    # form + the modular-add kernel, timed on the device
    exchange = None
    if world > 1:
        from pyfhesi.sharded import allgather_add
        part = torch.randint(0, 1 << 29, (14, 3, dev.Lt, dev.N), dtype=torch.int32, device="cuda")
        for _ in range(3):
            allgather_add(dev, part, 3 * 14)
        barrier()
        ev0.record(stream)
        reps = 20
        for _ in range(reps):
            allgather_add(dev, part, 3 * 14)
        ev1.record(stream)
        barrier()
        ex_ms = ev0.elapsed_time(ev1) / reps
        exchange = {"what": "ncclAllGather of 14 x 3 x Lt x N words per rank + k_tprod_reduce_world",
                    "bytes_per_rank": part.numel() * 4, "bytes_received_per_rank": part.numel() * 4 * (world - 1),
                    "us": ex_ms * 1e3}

    if not args.no_regression and world > 1:
        regression = run_regression_ours(args, rank, local_rank, world, lib_path)

    ms_step = ms_total / args.steps
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, pcie_ms, e2e_blocking_ms] + ([exchange["us"]] if exchange else []),
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        vals = t.tolist()
        ms_step, e2e_ms, pcie_ms, e2e_blocking_ms = vals[:4]
        if exchange:
            exchange["us"] = vals[4]
            exchange["algbw_GBs"] = exchange["bytes_received_per_rank"] / (vals[4] * 1e-6) / 1e9
    value = world * B / (ms_step * 1e-3)
    e2e_value = world * B / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_mont32 = dev.modmul_peak(32)
        peak64 = dev.modmul_peak(64)
        pipes = {name: dev.pipe_peak(kind) for kind, name in
                 enumerate(["imad_lo32", "imad_wide64", "imad_hi32", "shoup_modmul32", "alu_csub", "dfma"])}
        # The integer-multiply (FMA-heavy) pipe executes three instruction classes at three different
        # measured rates; a kernel's roofline is the pipe TIME its algorithmic instructions need.  Work is
        # therefore counted per class and expressed in Shoup-modmul equivalents (one Shoup product =
        # 2 IMAD + 1 IMAD.HI; one 64-bit multiply-accumulate = 1 IMAD.WIDE, about 0.66 of a Shoup product):
        # achieved / peak is then the fraction of the pipe's time spent on algorithmic instructions, the
        # quantity ncu reports as sm__inst_executed_pipe_fmaheavy (profiles/r02f_ncu_full_summary.txt).
        cost = {"lo": 1.0 / pipes["imad_lo32"], "hi": 1.0 / pipes["imad_hi32"], "wide": 1.0 / pipes["imad_wide64"]}
        shoup_cost = 2 * cost["lo"] + cost["hi"]
        peak32 = 1.0 / shoup_cost  # Shoup products per second when nothing else shares the pipe
        work_all = kernel_pipe_work_per_op(dev)
        pipe_s = lambda w: w["lo"] * cost["lo"] + w["hi"] * cost["hi"] + w["wide"] * cost["wide"]
        base = lambda k: {"k_residues_t": "k_residues", "k_fused_tensor_2k": "k_fused_tensor",  # template instances, N = 2048
                          "k_fused_keyswitch_split_2k": "k_fused_keyswitch_split"}.get(k.split("<")[0], k.split("<")[0])
        work = {k: work_all[base(k)] for k in prof if base(k) in work_all}
        top = max(prof.items(), key=lambda kv: kv[1][1]) if prof else (None, (0, 0.0))
        tname, (tcnt, tms) = top
        share = tms / max(sum(v[1] for v in prof.values()), 1e-9)
        ops_timed = B * args.steps
        tw = work.get(tname, {"lo": 0, "hi": 0, "wide": 0, "modmul": 0, "mac": 0})
        eq_per_launch = pipe_s(tw) / shoup_cost * ops_timed / max(tcnt, 1)
        avg_launch_s = tms * 1e-3 / max(tcnt, 1)
        achieved = eq_per_launch / max(avg_launch_s, 1e-12)
        canon = CANON_MODMUL[logq]
        step_pipe_s = sum(pipe_s(w) for w in work.values())
        traffic = kernel_compulsory_bytes_per_op(dev).get(base(tname or ""))
        roofline = {
            "bound": "integer-multiply (FMA-heavy) pipe; SURVEY.md §8d -- HBM does not bind",
            "kernel": tname, "kernel_share_of_step": share,
            "achieved": achieved / 1e9, "peak": peak32 / 1e9,
            "unit": "G Shoup-modmul equivalents/s (32-bit; every multiply-pipe instruction class weighted by its "
                    "measured cost: IMAD 1, IMAD.HI %.2f, IMAD.WIDE %.2f IMAD slots)" % (
                        cost["hi"] / cost["lo"], cost["wide"] / cost["lo"]),
            "frac": achieved / peak32,
            "frac_is": "pipe-time fraction: algorithmic multiply-pipe instructions of the kernel x their measured "
                       "per-class cost / the kernel's measured duration (comparable with ncu's fmaheavy pipe utilisation)",
            "work_per_op": {"shoup_modmul": tw["modmul"], "mac64": tw["mac"], "imad": tw["lo"], "imad_hi": tw["hi"],
                            "imad_wide": tw["wide"]},
            "traffic": traffic * ops_timed / max(tcnt, 1) if traffic else None,
            "traffic_unit": "bytes per launch: the kernel's compulsory HBM traffic (every input word read once, every "
                            "output word written once; key tiles and tables stay in L2), confirmed against ncu "
                            "dram__bytes_read.sum + dram__bytes_write.sum of the same launch in profiles/r02f_ncu_full_summary.txt",
            "peak_source": "measured in this run (fhesi_pipe_peak: register-resident ILP-8 chains of one instruction "
                           "class on all SMs)",
            "peak_montgomery32_Gmodmul_s": peak_mont32 / 1e9,
            "pipe_peaks_Gops_s": {k: v / 1e9 for k, v in pipes.items()},
            "avg_launch_ms": avg_launch_s * 1e3,
            "whole_op": {
                "pipe_frac": (value / world) * step_pipe_s,
                "shoup_modmul_eq_per_op": step_pipe_s / shoup_cost,
                "work_reduction_vs_reference_formulation": {
                    "canonical_modmul64_eq_per_op": canon,
                    "note": "SURVEY.md §8d counts the reference's formulation (64-bit products, Bluestein at 2N, "
                            "3D*L digit transforms); this build executes a cheaper one (30-bit limbs, zero-padded "
                            "2^k transform, split keys), so that count is a work figure, not a roofline numerator",
                    "peak64_Gmodmul_s": peak64 / 1e9}},
            "hbm": {"algorithmic_bytes_per_op": algorithmic_bytes_per_op(n, logq),
                    "achieved_GBs": (value / world) * algorithmic_bytes_per_op(n, logq) / 1e9,
                    "peak_GBs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "frac": (value / world) * algorithmic_bytes_per_op(n, logq) / 1e9 / hbm_peak,
                    "compulsory_bytes_per_op_all_kernels": sum(kernel_compulsory_bytes_per_op(dev).get(base(k), 0)
                                                               for k in prof)},
            "per_kernel_ms": {k: {"launches": v[0], "ms": v[1],
                                  "pipe_frac": (pipe_s(work[k]) * ops_timed / (v[1] * 1e-3)) if k in work and v[1] > 0 else None}
                              for k, v in sorted(prof.items())},
        }
        cpu = cpu_baseline_single_core(logq, p, g) if world == 1 and not args.no_cpu else None
        if regression is not None and world == 1 and not args.no_cpu:
            regression["cpu_baseline"] = regression_reference(full_tail=False)
        line = {
            "metric": METRIC, "value": value, "unit": "ops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": bench_config(B, world),
            "impl_details": {"chain": f"{dev.Lt} x 30-bit primes (tensor), {dev.Lk} (key switch"
                                      + (f", {dev.Ls} with split keys" if dev.Ls else "") + f"), N={dev.N}",
                             "host_numa_node_rank0": numa_node},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "ops/s", "h2d_bytes_per_step": 2 * B * ct_words * 4,
                    "d2h_bytes_per_step": B * ct_words * 4, "ms_per_step": e2e_ms,
                    "api": "fhesi_mult_relin_host_async, one call per step, fhesi_sync_all after the last (page-locked host "
                           "buffers" + ("; operands write-combined)" if wc else ")") + ": step i + 1 is enqueued while step i "
                           "drains, every step's operands come from host memory and every step's result lands in host memory "
                           "inside the timed region",
                    "blocking_call": {"value": world * B / (e2e_blocking_ms * 1e-3), "unit": "ops/s",
                                      "ms_per_step": e2e_blocking_ms,
                                      "api": "fhesi_mult_relin_host: returns when the step's result is in host memory"},
                    "host_copy_bound": {"ms_per_step": pcie_ms, "ops_s": world * B / (pcie_ms * 1e-3),
                                        "GBs_all_ranks": world * 3 * B * ct_words * 4 / (pcie_ms * 1e-3) / 1e9,
                                        "e2e_over_bound": e2e_value / (world * B / (pcie_ms * 1e-3)),
                                        "what": "the step's H2D + D2H bytes copied with no kernels, both directions "
                                                "at once, on all ranks concurrently (max over ranks)"}},
            "exchange": exchange,
            "regression": regression,
            "gpu_launches": launches,
            "crt_fallback_rate": crt_fallbacks / max(1.0, 3.0 * n * B * (args.steps + args.warmup)),
            "clocks": sampler.result(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_pipe_work_per_op(dev):
    """Integer-multiply-pipe instructions per mult+relin op, by kernel and instruction class, read off the
    kernel sources (DESIGN.md "work accounting"):
      Shoup product (butterfly twiddle, Garner step)   2 IMAD + 1 IMAD.HI
      Montgomery product of two residues               1 IMAD + 2 IMAD.WIDE
      64-bit multiply-accumulate                       1 IMAD.WIDE
      Montgomery reduction of a 64-bit sum             1 IMAD + 1 IMAD.WIDE   (k_residues: 1 IMAD + 1 IMAD.HI)
    A fused 1024-point transform is 36 Shoup products per thread x 128 threads (4608: the first forward stage
    sees a zero upper half, the last stage's twiddle is 1)."""
    N, n, D, W = dev.N, dev.n, dev.D, dev.W
    Lt, Lk, Ls = dev.Lt, dev.Lk, dev.Ls
    K = 3 * D
    garner = lambda L: L * (L - 1) // 2
    horner = lambda L: L * (L + 1) // 2
    tr = 4608 if N == 1024 else (10240 if N == 2048 else (N // 2) * int(math.log2(N)))  # fused: the last stage's twiddle is 1

    def w(modmul=0, montmul=0, mac=0, red=0, red_hi=0):
        return {"lo": 2 * modmul + montmul + red + red_hi, "hi": modmul + red_hi, "wide": 2 * montmul + mac + red,
                "modmul": modmul + montmul, "mac": mac}
    return {
        "k_residues": w(mac=4 * n * Lt * W, red_hi=4 * n * Lt * ((W + 3) // 4)),
        "k_fused_tensor": w(modmul=7 * Lt * tr, montmul=4 * Lt * N),
        "k_crt": w(modmul=3 * n * garner(Lt), mac=3 * n * horner(Lt)),
        # windowed explicit CRT: L Shoup products (y_i), 2 L multiply-accumulates for the quotient, (L + 1) NL for the
        # 28-bit limb columns, NL = limbs from two below the rounding bit's up to bit 2 logQ - 1
        "k_crt_direct": w(modmul=3 * n * Lt,
                          mac=3 * n * (2 * Lt + (Lt + 1) * ((2 * dev.info.logQ - 1) // 28 - max((dev.info.logQ - 1) // 28 - 2, 0) + 1))),
        "k_fused_keyswitch_split": w(modmul=(K + 4) * Ls * tr, mac=4 * K * Ls * N, red=4 * Ls * N),
        "k_fused_keyswitch": w(modmul=(K + 2) * Lk * tr, mac=2 * K * Lk * N, red=2 * Lk * N),
        "k_crt_split": w(modmul=4 * n * garner(Ls), mac=4 * n * horner(Ls)),
        # generic path (one butterfly per thread per stage, Montgomery twiddles)
        "k_fwd": w(montmul=(4 * Lt + K * Lk) * (N // 2) * int(math.log2(N))),
        "k_inv": w(montmul=(3 * Lt + 2 * Lk) * (N // 2) * int(math.log2(N))),
        "k_tensor_pw": w(montmul=4 * Lt * N),
        "k_dot": w(mac=2 * K * Lk * N),
    }


def kernel_compulsory_bytes_per_op(dev):
    """HBM bytes each hot kernel must move per op: inputs read once + outputs written once (tables and key
    tiles are shared by the whole grid and stay in L2)."""
    n, D, W, Lt, Ls, Lk = dev.n, dev.D, dev.W, dev.Lt, dev.Ls, dev.Lk
    K = 3 * D
    return {
        "k_residues": 4 * n * W * 4 + 4 * Lt * n * 4,
        "k_fused_tensor": 4 * Lt * n * 4 + 3 * Lt * n * 4,
        "k_crt": 3 * Lt * n * 4 + K * n * 4,
        "k_crt_direct": 3 * Lt * n * 4 + K * n * 4,
        "k_fused_keyswitch_split": K * n * 4 + 4 * Ls * n * 4,
        "k_fused_keyswitch": K * n * 4 + 2 * Lk * n * 4,
        "k_crt_split": 4 * Ls * n * 4 + 2 * n * W * 4,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="ciphertext pairs per GPU per step (SURVEY.md §8d: B in {1, 64, 1024, 8192})")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-regression", action="store_true", help="skip the Test_Regression leg (BASELINE metric 2)")
    ap.add_argument("--prime", type=int, default=1019, choices=[1019, 2027],
                    help="plaintext modulus p (m = p - 1); 2027 is the reference README's second family (N = 2048), "
                         "not a BASELINE configuration")
    ap.add_argument("--logq", type=int, default=256, choices=[128, 256, 512],
                    help="BASELINE config 5 sweep; the headline metric is quoted at 256")
    args = ap.parse_args()
    CFG["logQ"] = args.logq
    CFG["p"] = args.prime
    global METRIC
    METRIC = f"ciphertext mult+relin/sec at logQ={args.logq}"
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
