#!/usr/bin/env python
"""The reference's size sweep (run_simulations.sh:26-60) through the reference's OWN driver built with
BruteForce_CUDA (integration/build_patched_reference.py), restricted to the CUDA method (-m c):

    N in {1e3, 1e4, 1e5, 2e5, 5e5, 1e6, 2e6, 5e6} x D in {2, 3}, accuracy off   (run_simulations.sh:39-47)
    the first three sizes again with -a 1 -m c  (accuracy column against the reference's CPU brute force)

Each run is the unmodified CLI (main.cpp:885-927); the CSV row the driver writes
(`BruteForce_CUDA,n,D,seconds[,accuracy]`, main.cpp:59-63 schema) is collected into one table next to
the rows the reference's notebook pasted by hand (analysis/aggregated_results.csv:227-234: 2D only,
N=1e6: 8.64 s).  Times are the driver's own wall clock around the method call (safely_execute,
utils.h:87-104): host vectors in, host vectors out, H2D/D2H included.

    python integration/run_sweep.py [--out gpurun_out/sweep] [--max-n 5000000] [--precisions 64,32]
"""
import argparse
import csv
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "integration", "nbody_sim")
N_VALUES = [1000, 10000, 100000, 200000, 500000, 1000000, 2000000, 5000000]   # run_simulations.sh:26


def run_one(n, dim, accuracy, precision, timeout):
    work = tempfile.mkdtemp(prefix="nb200_sweep_")
    try:
        env = dict(os.environ, NB200_PRECISION=str(precision))
        r = subprocess.run([EXE, "-N", str(n), "-d", str(dim), "-a", str(accuracy), "-m", "c"], cwd=work, env=env,
                           capture_output=True, text=True, timeout=timeout)
        rows = []
        for f in glob.glob(os.path.join(work, "results", "*.csv")):
            for row in csv.reader(open(f)):
                if row and row[0] == "BruteForce_CUDA":
                    rows.append(row)
        if r.returncode != 0 or not rows:
            return {"n": n, "dim": dim, "precision": precision, "error": (r.stdout + r.stderr)[-400:]}
        row = rows[-1]
        t = float(row[3])
        out = {"method": row[0], "n": int(row[1]), "dim": int(row[2]), "seconds": t, "precision": precision,
               "G_interactions_per_s": round(n * (n - 1.0) / t / 1e9, 2) if t > 0 else None}
        if accuracy and len(row) > 4:
            out["accuracy_pct"] = float(row[4])
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep"))
    ap.add_argument("--max-n", type=int, default=5000000)
    ap.add_argument("--precisions", default="64,32")
    a = ap.parse_args()
    if not os.path.exists(EXE):
        sys.exit(f"{EXE} is missing: run integration/build_patched_reference.py where /root/reference exists")
    os.makedirs(a.out, exist_ok=True)
    results = []
    for precision in [int(p) for p in a.precisions.split(",")]:
        for dim in (2, 3):
            for n in N_VALUES:
                if n > a.max_n:
                    continue
                results.append(run_one(n, dim, 0, precision, 600))
                print(json.dumps(results[-1]), flush=True)
        for dim in (2, 3):
            for n in N_VALUES[:3]:
                rec = run_one(n, dim, 1, precision, 600)
                rec["with_accuracy"] = True
                results.append(rec)
                print(json.dumps(rec), flush=True)
    with open(os.path.join(a.out, "bruteforce_cuda_rows.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Method", "Bodies", "Dimension", "Time(s)", "Accuracy(%)", "Precision", "G_interactions_per_s"])
        for r in results:
            if "error" in r:
                continue
            w.writerow([r["method"], r["n"], r["dim"], f"{r['seconds']:.6f}", r.get("accuracy_pct", ""), r["precision"],
                        r["G_interactions_per_s"]])
    json.dump(results, open(os.path.join(a.out, "bruteforce_cuda_rows.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
