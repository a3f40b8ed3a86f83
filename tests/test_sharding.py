"""Host-side logic of the target-sharded multi-GPU path, on CPU: shard geometry and a
world_size-2 gloo run of the bootstrap + result-assembly plumbing (oracle stands in for the
GPU kernel so that the N>1 host path is covered without GPUs)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1000, 16384, 262144, 1048576])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_ranges_partition_the_bodies(pkg, n, world):
    from importlib import import_module
    dist = import_module(pkg.__name__ + ".distributed")
    edges = [dist.shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
        assert lo <= hi == lo2
    for lo, hi in edges:
        assert lo % 256 == 0 or lo == n          # shards start on a tile boundary


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch.distributed as dist
    import __graft_entry__ as entry
    pkg = entry.load_package(); oracle = entry.load_oracle()
    from importlib import import_module
    D = import_module(pkg.__name__ + ".distributed")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. bootstrap: a 128-byte id made on rank 0 reaches every rank unchanged
    blob = bytes(range(128)) if rank == 0 else None
    got = D.broadcast_bytes(blob, 128, src=0)
    assert got == bytes(range(128)), "unique-id broadcast corrupted"
    # 2. shard the targets, compute own rows (oracle as stand-in kernel), assemble on every rank
    n = 1000
    bodies = pkg.generators.uniform_cube(n, 3, seed=21)
    lo, hi = D.shard_range(n, rank, world)
    rows = np.zeros((n, 3))
    rows[lo:hi] = oracle.forces_targets(bodies, np.arange(lo, hi))
    full = D.assemble_rows(rows, lo, hi)
    assert np.array_equal(full, oracle.forces(bodies)), "sharded rows != full force evaluation"
    # 3. one sharded step == one full step: integrate own rows, all-gather positions
    dt = 1e-3
    mine = bodies.copy()
    mine[lo:hi, 3:6] += full[lo:hi] / mine[lo:hi, 6:7] * dt
    mine[lo:hi, 0:3] += mine[lo:hi, 3:6] * dt
    stepped = D.assemble_rows(mine, lo, hi)
    assert np.array_equal(stepped, oracle.simulate(bodies, dt, 1)), "sharded step != full step"
    dist.barrier()
    if rank == 0: print("GLOO_OK")
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_OK" in r.stdout


def test_stratified_generator_keeps_bodies_apart(pkg):
    for dim, n in ((2, 4096), (3, 1000), (2, 1000)):
        b = pkg.generators.jittered_cube(n, dim, seed=3)
        assert b.shape == (n, 2 * dim + 1) and b[:, :dim].min() >= 0.0 and b[:, :dim].max() <= 1.0
        k = int(np.ceil(n ** (1.0 / dim) - 1e-9))
        d = b[:, None, :dim] - b[None, :512, :dim]
        r = np.sqrt((d * d).sum(-1))
        r[r == 0] = np.inf
        assert r.min() >= 0.5 / k - 1e-12


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("n", [1, 1000, 16384, 29000, 262144, 1 << 20])
def test_pair_symmetric_work_lists_cover_every_pair_once(pkg, lib, n, world):
    """Host logic of the cross-rank pair-symmetric pass (nb_force_sym.cuh / build_sym_rows): over all
    ranks, every ordered (target tile, source tile) pair must be delivered exactly once -- by an
    ordered row, or by a symmetric row that serves both directions -- and every rank must get the
    same share of the work."""
    import ctypes
    tile, tpi = 256, 4
    tiles = max(1, -(-n // tile))
    T = -(-tiles // world)
    NT = T * world
    cover = np.zeros((NT, NT), dtype=np.int32)         # [target tile, source tile] deliveries
    work = []
    for rank in range(world):
        cap = 8 * (T // tpi + 2)
        buf = (ctypes.c_int * (4 * cap))()
        nrows = lib.nb200_debug_sym_rows(n, world, rank, buf, cap)
        assert 0 < nrows <= cap
        rows = np.frombuffer(buf, dtype=np.int32)[:4 * nrows].reshape(nrows, 4)
        w = 0
        for it, t0, t1, flags in rows:
            i0 = rank * T + it * tpi
            i1 = min(i0 + tpi, (rank + 1) * T)         # lanes past the shard are inert
            assert 0 <= t0 < t1 <= NT and i0 < i1
            cover[i0:i1, t0:t1] += 1
            if flags & 1:
                cover[t0:t1, i0:i1] += 1               # the reaction: sources become targets
                assert t0 >= i1 or t1 <= i0            # a symmetric row never holds its own i-tile
            w += (i1 - i0) * (t1 - t0)
        work.append(w)
    assert cover.min() == 1 and cover.max() == 1
    if T % (2 * tpi) == 0:                             # whole i-tiles and an even split of the opposite block
        assert max(work) == min(work)
    else:
        assert max(work) - min(work) <= 2 * tpi * T


def test_device_generator_mirror_known_answers(pkg):
    """The numpy mirror of the device generator: Philox-4x32-10 known-answer vector (Random123:
    counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8) and the ranges of utils.h:113-115."""
    g = pkg.generators
    a, b = g._philox_uniforms(0, np.array([0], dtype=np.uint64), 0)
    assert a[0] == (0x6627E8D5E169C58D >> 11) * 2.0 ** -53 and b[0] == (0xBC57AC4C9B00DBD8 >> 11) * 2.0 ** -53
    r = g.device_bodies(20000, 3, 0, seed=3)
    assert 1.0 <= r[:, :3].min() and r[:, :3].max() <= 1.0e7 and np.abs(r[:, 3:6]).max() <= 10.0
    assert 1.0 <= r[:, 6].min() and r[:, 6].max() <= 1.0e8
    u = g.device_bodies(20000, 2, 1, seed=3)
    assert 0.0 <= u[:, :2].min() and u[:, :2].max() < 1.0 and abs(u[:, :2].mean() - 0.5) < 0.01
    p = g.device_bodies(50000, 3, 2, seed=3)
    rad = np.linalg.norm(p[:, :3], axis=1)
    assert abs(np.median(rad) - 1.3048) < 0.02 and rad.max() < 22.8     # Plummer half-mass radius
    assert not np.array_equal(g.device_bodies(100, 3, 1, seed=1), g.device_bodies(100, 3, 1, seed=2))


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
def test_reaction_sum_exchange_is_consistent_across_ranks(pkg, lib, world):
    """Cross-rank pair-symmetric pass: rank g pushes the reaction sums on rank h's bodies into slot k
    of h exactly when h expects rank g in its slot k; every rank whose bodies g evaluated as SOURCES
    (symmetric rows of its work list) receives a push; slots of one receiver never collide."""
    import ctypes
    plans = []
    for rank in range(world):
        buf = (ctypes.c_int * 24)()
        k = lib.nb200_debug_sym_exchange(world, rank, buf, 8)
        assert k == world // 2
        plans.append([(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(k)])
    for g, plan in enumerate(plans):
        for send_to, recv_from, slot in plan:
            assert send_to != g and recv_from != g and 0 <= slot < world // 2
            assert (g, slot) in [(rf, sl) for _, rf, sl in plans[send_to]]      # receiver expects g in that slot
            assert (g, slot) in [(st, sl) for st, _, sl in plans[recv_from]]    # the expected sender targets that slot
        assert len({sl for _, _, sl in plan}) == len(plan)
    # every shard a rank touches as symmetric SOURCES is one it pushes to
    n = 8192 * world
    T = -(-(-(-n // 256)) // world)
    for rank in range(world):
        buf = (ctypes.c_int * (4 * 4096))()
        nrows = lib.nb200_debug_sym_rows(n, world, rank, buf, 4096)
        rows = np.frombuffer(buf, dtype=np.int32)[:4 * nrows].reshape(nrows, 4)
        touched = {int(t0) // T for _, t0, _, fl in rows if fl & 1} | {int(t1 - 1) // T for _, _, t1, fl in rows if fl & 1}
        touched.discard(rank)
        assert touched <= {st for st, _, _ in plans[rank]}
