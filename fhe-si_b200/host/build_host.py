"""Build libfhesi_host.so: the reference-named C++ classes over the C ABI (g++, no CUDA code).
`backend` is the library that provides the fhesi_* symbols: libfhesi_b200.so (product) or, for
GPU-less CI only, the kernel-logic emulator build under tests/emu/_build."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_host(backend, out_dir=None, force=False):
    out_dir = out_dir or os.path.dirname(backend)
    out = os.path.join(out_dir, "libfhesi_host.so")
    deps = [os.path.join(HERE, f) for f in ("fhesi_host.cpp", "fhesi_host.h", "ntl_shim.h")]
    deps.append(os.path.join(ROOT, "include", "fhesi.h"))
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    import sys
    sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200"))
    from buildlock import build_lock, publish
    fresh = lambda: os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps)
    bdir, bname = os.path.dirname(backend), os.path.basename(backend)
    assert bname.startswith("lib") and bname.endswith(".so")
    with build_lock(out):  # several ranks may get here at once: one builds, the others find it fresh
        if not force and fresh():
            return out
        tmp = out + ".tmp.%d" % os.getpid()
        cmd = ["g++", "-std=c++17", "-O3", "-g", "-fPIC", "-shared", "-pthread", "-Wall", "-Wno-sign-compare",
               "-I", HERE, "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "fhesi_host.cpp"),
               "-o", tmp, "-L", bdir, "-l" + bname[3:-3], "-Wl,-rpath," + bdir]
        subprocess.check_call(cmd)
        publish(tmp, out)
    return out


def compile_client(sources, backend, out, extra_includes=(), defines=()):
    """Compile a client program (sources written against the reference's headers) against the
    host layer.  extra_includes are searched AFTER fhe-si_b200/host, so same-named reference
    headers are shadowed by ours while client-only files (Matrix.*, Regression.h) are found."""
    host = build_host(backend)
    hdir, bdir = os.path.dirname(host), os.path.dirname(backend)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-w", "-I", HERE, "-I", os.path.join(ROOT, "include")]
    for inc in extra_includes:
        cmd += ["-I", inc]
    cmd += ["-D" + d for d in defines] + list(sources) + ["-o", out, "-L", hdir, "-lfhesi_host", "-L", bdir,
            "-l" + os.path.basename(backend)[3:-3], "-Wl,-rpath," + hdir, "-Wl,-rpath," + bdir, "-pthread"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    import sys
    print(build_host(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fhe-si_b200", "libfhesi_b200.so"), force=True))
