#!/usr/bin/env python
"""torchrun worker: one rank per GPU, rank contexts with the fused NVLink exchange (or NCCL), checks
the sharded multi-step trajectory against a 1-GPU run of the same library on rank 0.

    python -m torch.distributed.run --nproc-per-node 2 tools/mp_check.py [--exchange p2p|nccl] [--overlap 0|1]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from importlib import import_module

    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default="auto")
    ap.add_argument("--overlap", type=int, default=1)
    ap.add_argument("--n", type=int, default=30000)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--precision", type=int, default=64)
    ap.add_argument("--detect", type=int, default=-1, help="1 = force the close-pair pre-pass (FP32: cross-rank pair-symmetric pass)")
    a = ap.parse_args()
    pkg = entry.load_package()
    D = import_module(pkg.__name__ + ".distributed")
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bodies = pkg.generators.plummer(a.n, seed=11)
    ctx = D.create_rank_context(pkg, 3, a.n, a.precision, device=local, exchange=a.exchange)
    ctx.set_option("overlap", a.overlap)
    if a.detect >= 0:
        ctx.set_option("detect", a.detect)
    lo, hi = ctx.shard_range()
    assert (lo, hi) == D.shard_range(a.n, rank, world)
    out = bodies.copy()
    # two epochs: upload -> steps -> download, twice (exercises the epoch handshake of re-uploads)
    for epoch in range(2):
        ctx.upload(out)
        ctx.step(1e-3, a.steps // 2)
        ctx.step(1e-3, a.steps - a.steps // 2)
        mine = out.copy()
        ctx.download(mine)
        out = D.assemble_rows(mine, lo, hi)
    forces = np.zeros((a.n, 3))
    ctx.upload(out)
    ctx.forces(out=forces)
    forces = D.assemble_rows(forces, lo, hi)
    plan = ctx.plan
    ctx.close()
    ok = True
    if rank == 0:
        one = bodies.copy()
        with pkg.NBodyCuda(3, a.n, a.precision) as c1:
            if a.detect >= 0:
                c1.set_option("detect", a.detect)
                c1.set_option("symmetric", 0)
            for epoch in range(2):
                c1.upload(one)
                c1.step(1e-3, a.steps)
                c1.download(one)
            c1.upload(one)
            f1 = c1.forces()
        ex = np.abs(out[:, :6] - one[:, :6]).max() / np.abs(one[:, :6]).max()
        eff = pkg.generators.relative_norm_error(forces, f1)
        if a.precision == 64:
            ef, tol, ftol = eff.max(), 1e-10, 1e-10
        else:
            # FP32 sums taken in another order (pair-symmetric vs ordered pass): the trajectories part at
            # the 1e-7 level and forces of close pairs amplify that, so hold the bulk, not the worst body
            ef, tol, ftol = float(np.percentile(eff, 99)), 1e-5, 1e-4
        ok = bool(ex <= tol and ef <= ftol)
        print(f"MP_CHECK world={world} exchange={a.exchange} overlap={a.overlap} traj_err={ex:.3e} force_err={ef:.3e} "
              f"{'OK' if ok else 'FAIL'} | {plan}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
