#!/usr/bin/env python
"""Small pair-symmetric + ordered runs for compute-sanitizer (memcheck / racecheck), one tool per call:

    compute-sanitizer --tool racecheck python tools/sanitize_case.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
for dim, n in ((3, 2600), (2, 1500)):
    b = pkg.generators.uniform_cube(n, dim, seed=2)
    for prec in (32, 64):
        for opts in ({"detect": 1, "symmetric": 1}, {"detect": 1, "symmetric": 1, "sym_ti": 8}, {"detect": 0}):
            with pkg.NBodyCuda(dim, n, prec) as ctx:
                for k, v in opts.items():
                    ctx.set_option(k, v)
                ctx.upload(b)
                f = ctx.forces()
                ctx.step(1e-5, 2)
                assert np.all(np.isfinite(f))
print("sanitize_case done")
