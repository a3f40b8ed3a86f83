#!/usr/bin/env python
"""Sweep rows of the patched suite copy -> the layout of the reference's analysis/aggregated_results.csv
(`Bodies,Method,Dimension,Average Runtime (s)`, lines 227-234 hold its hand-pasted 2D BruteForce_CUDA rows), so that
analyze_results.ipynb (cell 4 reads that file) plots measured rows for 2D and 3D without change.

    python integration/aggregate_rows.py profiles/r02/sweep/bruteforce_cuda_rows_fp64.csv > rows.csv
Rows of the same (N, D) -- the sweep runs the first four sizes twice, once with -a 1 -- are averaged like the notebook does."""
import csv
import sys
from collections import defaultdict


def main():
    acc = defaultdict(list)
    for path in sys.argv[1:]:
        with open(path) as f:
            for row in csv.DictReader(f):
                acc[(int(row["Bodies"]), row["Method"], int(row["Dimension"]))].append(float(row["Time(s)"]))
    w = csv.writer(sys.stdout, lineterminator="\n")
    w.writerow(["Bodies", "Method", "Dimension", "Average Runtime (s)"])
    for (n, method, dim), ts in sorted(acc.items(), key=lambda kv: (kv[0][1], kv[0][2], kv[0][0])):
        w.writerow([n, method, dim, repr(sum(ts) / len(ts))])


if __name__ == "__main__":
    main()
