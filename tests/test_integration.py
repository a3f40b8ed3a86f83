"""The C++ host side: the methods.h-style adapter (host/methods_cuda.h) and the reference's own
benchmark driver rebuilt with BruteForce_CUDA as a selectable method (integration/)."""
import csv
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "build", "test_adapter")
NBODY_SIM = os.path.join(ROOT, "build", "integration", "nbody_sim")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_adapter_fails_loudly_without_gpu():
    """No CPU fallback: without a device the adapter throws (safely_execute would log and skip)."""
    if not os.path.exists(ADAPTER):
        pytest.skip("build/test_adapter not built")
    if _have_gpu():
        pytest.skip("GPU present")
    r = subprocess.run([ADAPTER, "3", "64"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("precision", ["64", "32"])
def test_cpp_adapter_matches_oracle(dim, precision):
    if not os.path.exists(ADAPTER):
        pytest.skip("build/test_adapter not built")
    env = dict(os.environ, NB200_PRECISION=precision)
    r = subprocess.run([ADAPTER, str(dim), "3000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "ADAPTER_OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("dim", [2, 3])
def test_reference_driver_emits_bruteforce_cuda_row(tmp_path, dim):
    """`nbody_sim -m ac -a 1`: the reference's main.cpp (patched copy) runs its own CPU variants and
    BruteForce_CUDA side by side; the -a 1 column (utils.h:170-219 vs brute_force_seq) must be 100 %."""
    if not os.path.exists(NBODY_SIM):
        pytest.skip("build/integration/nbody_sim not built (needs /root/reference at build time)")
    r = subprocess.run([NBODY_SIM, "-N", "3000", "-d", str(dim), "-a", "1", "-m", "ac"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    files = glob.glob(os.path.join(tmp_path, "results", "*.csv"))
    assert len(files) == 1
    rows = {row["Method"]: row for row in csv.DictReader(open(files[0]))}
    assert "BruteForce_CUDA" in rows, r.stdout[-3000:]
    assert rows["BruteForce_CUDA"]["Bodies"] == "3000" and rows["BruteForce_CUDA"]["Dimension"] == str(dim)
    assert float(rows["BruteForce_CUDA"]["Accuracy(%)"]) == 100.0
    assert float(rows["BruteForce_CUDA"]["Time(s)"]) > 0
    for peer in ("BruteForce_Sequential", "BruteForce_OpenMP2", "BruteForce_Parlay2"):
        assert peer in rows


@pytest.mark.gpu
def test_reference_driver_default_method_set_includes_cuda(tmp_path):
    """run_simulations.sh never passes -m (run_simulations.sh:16): the default set must include it."""
    if not os.path.exists(NBODY_SIM):
        pytest.skip("build/integration/nbody_sim not built")
    env = dict(os.environ, NB200_PRECISION="32")
    r = subprocess.run([NBODY_SIM, "-N", "2000", "-d", "3", "-m", "c"], cwd=tmp_path, capture_output=True, text=True,
                       timeout=300, env=env)
    assert r.returncode == 0
    rows = [row for f in glob.glob(os.path.join(tmp_path, "results", "*.csv")) for row in csv.DictReader(open(f))]
    assert [row["Method"] for row in rows] == ["BruteForce_CUDA"]
