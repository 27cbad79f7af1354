"""Inter-process lock + atomic publish for the in-tree builds.  Several ranks of one job (torchrun)
import the package at the same moment; if an artefact is stale each of them would otherwise
recompile it into the same path and load half-written files."""
import contextlib
import fcntl
import os


@contextlib.contextmanager
def build_lock(target):
    os.makedirs(os.path.dirname(target), exist_ok=True)
    with open(target + ".lock", "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def publish(tmp, target):
    """Atomically move a finished build product into place."""
    os.replace(tmp, target)
