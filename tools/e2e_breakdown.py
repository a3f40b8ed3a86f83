#!/usr/bin/env python
"""Wall-clock breakdown of the host-buffer path (upload / step / download) at one size.

    python tools/e2e_breakdown.py [--n 1048576] [--dim 3] [--precision 32] [--pageable]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--pageable", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    pkg = entry.load_package()
    bodies = pkg.generators.uniform_cube(a.n, a.dim, seed=1)
    host = bodies.copy() if a.pageable else torch.from_numpy(bodies.copy()).pin_memory().numpy()
    with pkg.NBodyCuda(a.dim, a.n, a.precision) as ctx:
        ctx.upload(host)
        ctx.step(1e-4, 1)
        ctx.download(host)
        t = {"upload_ms": [], "step_ms": [], "step_device_ms": [], "download_ms": []}
        for _ in range(a.reps):
            t0 = time.perf_counter(); ctx.upload(host)
            t1 = time.perf_counter(); ctx.step(1e-4, 1)
            t2 = time.perf_counter(); ctx.download(host)
            t3 = time.perf_counter()
            t["upload_ms"].append((t1 - t0) * 1e3); t["step_ms"].append((t2 - t1) * 1e3)
            t["step_device_ms"].append(ctx.last_elapsed_ms); t["download_ms"].append((t3 - t2) * 1e3)
    print(json.dumps({"n": a.n, "dim": a.dim, "precision": a.precision, "host": "pageable" if a.pageable else "pinned",
                      **{k: round(min(v), 3) for k, v in t.items()}}))


if __name__ == "__main__":
    main()
