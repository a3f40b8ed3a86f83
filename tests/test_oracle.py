"""The oracle (oracle/nbody_oracle.c) against the golden vectors generated from the reference's
own compiled methods.cpp (oracle/gen_golden.py), and against oracle/_ref directly where built."""
import numpy as np
import pytest

from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_forces_bit_exact_vs_golden(oracle, name):
    g = load_golden(name)
    G, cut = float(g["G"]), float(g["cutoff"])
    assert G == oracle.G_REF and cut == oracle.CUTOFF_REF
    # bit-exact: same operation order as methods.cpp:7-42 and :98-136
    assert np.array_equal(oracle.forces(g["bodies"], G, cut, "seq"), g["forces_seq"])
    assert np.array_equal(oracle.forces(g["bodies"], G, cut, "omp_2"), g["forces_omp2"])


@pytest.mark.parametrize("name", golden_names())
def test_simulate_bit_exact_vs_golden(oracle, name):
    g = load_golden(name)
    dt, ns = float(g["dt"]), int(g["nsteps"])
    assert np.array_equal(oracle.simulate(g["bodies"], dt, ns, variant="seq"), g["after_seq"])
    assert np.array_equal(oracle.simulate(g["bodies"], dt, ns, variant="omp_2"), g["after_omp2"])


@pytest.mark.parametrize("name", golden_names())
def test_variants_agree_normwise(oracle, pkg, name):
    """The reference's own variants differ by re-association only."""
    g = load_golden(name)
    rel = pkg.generators.relative_norm_error
    for k in ("forces_seq", "forces_omp1"):
        e = rel(g[k], g["forces_omp2"])
        assert np.all(np.isfinite(e)) and e.max() < 1e-12


def test_degenerate_semantics(oracle):
    """Hard cut-off (methods.cpp:119): duplicates and r^2 < 1e-10 pairs are skipped, r^2 = 1.21e-10 kept."""
    g = load_golden("degenerate3d_n9")
    b, f = g["bodies"], g["forces_omp2"]
    assert np.all(np.isfinite(f))
    # bodies 4,5 are 1.1e-5 apart -> kept: they dominate each other's force and are opposite
    f4, f5 = f[4], f[5]
    assert abs(f4[1]) > 1e30 and abs(f5[1]) > 1e30 and np.sign(f4[1]) == -np.sign(f5[1])
    # removing body 1 (the duplicate of body 0) must not change body 0's force from body 1: it was skipped
    keep = [i for i in range(len(b)) if i != 1]
    f_wo = oracle.forces(b[keep])
    assert np.allclose(f_wo[0], f[0], rtol=1e-13, atol=0)
    # likewise the r^2 = 0.81e-10 pair (2,3): dropping body 3 leaves body 2's force unchanged
    keep = [i for i in range(len(b)) if i != 3]
    assert np.allclose(oracle.forces(b[keep])[2], f[2], rtol=1e-13, atol=0)


def test_targets_subset_and_long_double(oracle, pkg):
    b = pkg.generators.uniform_cube(700, 3, seed=3)
    full = oracle.forces(b)
    idx = np.arange(0, 700, 13)
    assert np.array_equal(oracle.forces_targets(b, idx), full[idx])
    ld = oracle.forces_targets(b, idx, long_double=True)
    assert pkg.generators.relative_norm_error(full[idx], ld).max() < 1e-13


def test_energy_and_accuracy_metric(oracle, pkg):
    b = pkg.generators.uniform_cube(300, 2, seed=5)
    ke, pe = oracle.energy(b)
    m, v = b[:, 4], b[:, 2:4]
    assert np.isclose(ke, 0.5 * np.sum(m * np.sum(v * v, axis=1)), rtol=1e-13)
    d = b[:, None, :2] - b[None, :, :2]
    r2 = np.sum(d * d, axis=2)
    np.fill_diagonal(r2, np.inf)
    want = 0.25 * oracle.G_REF * np.sum(m[:, None] * m[None, :] / r2)
    assert np.isclose(pe, want, rtol=1e-12)
    f = oracle.forces(b)
    assert oracle.accuracy_pct(f, f) == 100.0
    bad = f.copy()
    bad[:30, 0] *= 1.02   # 2 % off on 30 bodies -> 10 % inaccurate (utils.h:170-219)
    assert np.isclose(oracle.accuracy_pct(bad, f), 90.0)


def test_energy_drift_is_small_for_small_dt(oracle, pkg):
    b = pkg.generators.plummer(256, seed=9)
    e0 = sum(oracle.energy(b))
    after = oracle.simulate(b, 1e-4, 20)
    e1 = sum(oracle.energy(after))
    assert abs(e1 - e0) / abs(e0) < 1e-3


def test_oracle_matches_compiled_reference_when_present(oracle, pkg):
    """In the build container oracle/_ref is the reference's own methods.cpp: require bit equality."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libnbref.so not present")
    for dim, n, seed in ((3, 513, 11), (2, 400, 12)):
        b = pkg.generators.reference_range(n, dim, seed)
        for v in ("seq", "omp_2"):
            assert np.array_equal(oracle.forces(b, variant=v), oracle.ref_forces(b, v)[0])
        assert np.array_equal(oracle.ref_forces(b, "parlay_2")[0], oracle.forces(b))
        assert np.array_equal(oracle.simulate(b, 0.5, 3), oracle.ref_simulate(b, 0.5, 3))
        f = oracle.forces(b)
        assert oracle.ref_accuracy_pct(f, f) == oracle.accuracy_pct(f, f) == 100.0
