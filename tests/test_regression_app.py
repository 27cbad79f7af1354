"""apps/regression_sharded.py end to end (BASELINE config 4 at toy size): encrypted regression
with blocks sharded over 2 gloo ranks, kernels in the test emulator; the decrypted theta*det and
det must equal the plaintext computation mod p.  The GPU run uses the same script over NCCL."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def run_app(extra, nproc, port=None, app="regression_sharded.py"):
    port = free_port()  # an unused rendezvous port: fixed numbers collide with lingering sockets
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "apps", app)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_sharded_regression_world2_emu(emu_lib):
    out = run_app(["--dim", "2", "--points", "20", "--prime", "23", "--gen", "7", "--lib", emu_lib, "--cpu-tensors"],
                  2, 29600 + os.getpid() % 300)
    assert out["correct"] and out["n_gpus"] == 2 and out["config"]["blocks"] == 3
    assert out["theta_det"] == out["expected"]


def test_regression_dim4_level_batched_adjugate_emu(emu_lib):
    """d = 4: the adjugate goes through 1x1, 2x2 and 3x3 minors, each level batched (apps/fhesi_app.py
    adjugate_and_det_batched); theta * det and det must equal RegressPT mod p."""
    out = run_app(["--dim", "4", "--points", "30", "--prime", "23", "--gen", "7", "--lib", emu_lib, "--cpu-tensors"], 1)
    assert out["correct"] and out["theta_det"] == out["expected"], out


def test_sharded_regression_from_shard_files_world2_emu(emu_lib, tmp_path):
    """README:82-84: the data split into files by the generator; each rank takes whole files and cuts
    each into its own blocks (3 files of 17, 17, 16 points, blocks of 8 -> 3 + 3 + 2 blocks, the last of each file ragged or full;
    the same 50 points as one set make 7 blocks)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from generate_random_data import main
    assert main(["x", str(tmp_path / "reg"), "2", "50", "3", "--seed", "5"]) == 0
    out = run_app(["--data", str(tmp_path / "reg"), "--prime", "23", "--gen", "7", "--lib", emu_lib, "--cpu-tensors"], 2)
    assert out["correct"] and out["n_gpus"] == 2 and out["config"]["blocks"] == 8, out
    assert out["theta_det"] == out["expected"] and "3 shard files" in out["config"]["input"]
    # the same data as one in-process set: same answer (the sums do not depend on the blocking)
    one = run_app(["--dim", "2", "--points", "50", "--seed", "5", "--prime", "23", "--gen", "7", "--lib", emu_lib,
                   "--cpu-tensors"], 1)
    assert one["theta_det"] == out["theta_det"] and one["config"]["blocks"] == 7


def test_sharded_statistics_world2_emu(emu_lib):
    """BASELINE config 3 at toy size: encrypted mean / covariance over 2 gloo ranks."""
    out = run_app(["--dim", "2", "--points", "20", "--prime", "23", "--gen", "7", "--lib", emu_lib, "--cpu-tensors"],
                  2, 29900 + os.getpid() % 90, app="statistics_sharded.py")
    assert out["correct"] and out["n_gpus"] == 2
    assert out["mean"] == out["expected"]["mean"] and out["cov_upper"] == out["expected"]["cov_upper"]
    assert out["N"] == 20 and out["N2"] == out["expected"]["N2"]


def test_data_generator_is_seeded_and_formatted(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from generate_random_data import generate, main
    a, b = generate(4, 50, 7), generate(4, 50, 7)
    assert a == b and len(a[0]) == 50 and all(-100 <= v <= 100 for r in a[0] for v in r)
    assert main(["x", str(tmp_path / "d"), "4", "50", "8", "--seed", "7"]) == 0
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 8  # README:82-84 split
    first = open(tmp_path / files[0]).read().splitlines()
    assert first[0] == "4 7" and len(first) == 8 and len(first[1].split()) == 5


@pytest.mark.gpu
def test_statistics_cfg3_gpu(cuda_lib):
    out = run_app(["--dim", "4", "--points", "10000"], 1, 29800 + os.getpid() % 90, app="statistics_sharded.py")
    assert out["correct"] and out["config"]["logQ"] == 100, out


@pytest.mark.gpu
def test_regression_cfg_small_gpu(cuda_lib):
    out = run_app(["--dim", "3", "--points", "2000"], 1, 29700 + os.getpid() % 200)
    assert out["correct"], out


@pytest.mark.gpu
def test_regression_cfg4_full_size_gpu(cuda_lib, tmp_path):
    """BASELINE config 4 at full size: d=4, N=100000 (seed 12345) split into 8 files, p=1019, g=3; parameters
    from the global N (logQ=176).  theta*det and det are the values the reference's RegressPT gives mod 1019
    (also what oracle/_ref's own Regression.h decrypts on a prefix of the data, tests/golden)."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from generate_random_data import main
    assert main(["x", str(tmp_path / "reg4"), "4", "100000", "8"]) == 0
    out = run_app(["--data", str(tmp_path / "reg4")], 1)
    assert out["correct"] and out["config"]["logQ"] == 176 and out["config"]["blocks"] == 392, out
    assert out["theta_det"] == [15, 282, 867, 160, 436], out
    gen = run_app(["--dim", "4", "--points", "100000"], 1)
    assert gen["theta_det"] == [15, 282, 867, 160, 436] and gen["config"]["blocks"] == 391, gen
