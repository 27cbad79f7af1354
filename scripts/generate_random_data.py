#!/usr/bin/env python3
"""Seeded Python 3 restatement of the reference's scripts/generateRandomData.py (Python 2,
unseeded): line 1 "d N", then N lines of d integers uniform in [-100, 100] and an integer label
= int(sum coeff_i x_i + gauss(0, 100)) with coeff_i ~ U(-10, 10); optional split into
name_k.dat files (README:82-84).

    python scripts/generate_random_data.py NAME d N [nFiles] [--seed S]
"""
import math
import random
import sys


def generate(d, n, seed=12345):
    """-> (rows [n][d], labels [n]) exactly as the files would hold them."""
    rnd = random.Random(seed)
    coeff = [rnd.uniform(-10, 10) for _ in range(d)]
    rows, labels = [], []
    for _ in range(n):
        val = [rnd.randint(-100, 100) for _ in range(d)]
        label = sum(c * v for c, v in zip(coeff, val)) + rnd.gauss(0, 100)
        rows.append(val)
        labels.append(int(label))  # '%d' % float truncates toward zero
    return rows, labels


def main(argv):
    seed = 12345
    if "--seed" in argv:
        i = argv.index("--seed")
        seed = int(argv[i + 1])
        del argv[i:i + 2]
    if len(argv) < 4:
        print("usage: generate_random_data.py filename d N [nFiles] [--seed S]")
        return 1
    name, d, n = argv[1], int(argv[2]), int(argv[3])
    nfiles = int(argv[4]) if len(argv) > 4 else 1
    rows, labels = generate(d, n, seed)
    per = int(math.ceil(float(n) / nfiles))
    for k in range(nfiles):
        lo, hi = k * per, min(n, (k + 1) * per)
        path = "%s_%d.dat" % (name, k) if nfiles > 1 else "%s.dat" % name
        with open(path, "w") as f:
            f.write("%d %d\n" % (d, hi - lo))
            for r, l in zip(rows[lo:hi], labels[lo:hi]):
                f.write(" ".join(str(v) for v in r) + " %d\n" % l)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
