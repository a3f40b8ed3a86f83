import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e
pkg = e.load_package(); oracle = e.load_oracle(); gen = pkg.generators
dim, n = 2, 20000
b = gen.round_to_float(gen.uniform_cube(n, dim, seed=44))
ref = oracle.forces(b); kappa = oracle.condition(b)
for opts in ({"symmetric": 0}, {"symmetric": 1, "detect": 0}, {"symmetric": 1, "detect": 1}, {"symmetric": 1, "detect": 0, "sym_ti": 4, "sym_block": 256}, {"symmetric": 1, "detect": 0, "sym_algo": 0, "sym_ti": 4, "sym_block": 256}):
    f = pkg.brute_force_cuda_n_body(b, 32, options=opts)
    err = gen.relative_norm_error(f, ref)
    bound = np.maximum(1e-5, 6e-7 * kappa)
    w = np.argmax(err / bound)
    print(opts, "worst ratio %.2f err %.3e kappa %.1f body %d; p99 %.2e; n over bound %d; err/kappa max %.2e" % (err[w]/bound[w], err[w], kappa[w], w, np.percentile(err, 99), (err > bound).sum(), (err/kappa).max()))
