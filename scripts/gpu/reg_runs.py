"""Ten runs of the sharded regression app (BASELINE config 4, one GPU) in one process, phases of every run: which part
of the clock is slow while a fresh box settles."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ("apps", "fhe-si_b200", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import bench  # noqa: E402
import build as fhesi_build  # noqa: E402
import regression_sharded  # noqa: E402
import tempfile

tmp = tempfile.mkdtemp(prefix="fhesi_reg_")
bench._regression_files(tmp)
R = bench.REG
a = types.SimpleNamespace(d=R["d"], n=R["n"], p=R["p"], g=R["g"], seed=R["seed"], data=os.path.join(tmp, "reg4"),
                          lib=fhesi_build.build(), cpu_tensors=False)
for i in range(10):
    r = regression_sharded.run(a, 0, 1, 0, quiet=True)
    print(i, round(r["value"], 4), {k: round(v, 4) for k, v in r["phases_s"].items()}, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["setup_split_s"].items() if k != "note"}, flush=True)
