"""The C++ host side: the methods.h-style adapter (host/methods_cuda.h) and the reference's own
benchmark driver rebuilt with BruteForce_CUDA as a selectable method (integration/)."""
import csv
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "build", "test_adapter")
NBODY_SIM = os.path.join(ROOT, "build", "integration", "nbody_sim")
SWEEP = os.path.join(ROOT, "build", "integration", "sweep")
REF = "/root/reference/nbody-sim-new"


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
@pytest.mark.parametrize("precision", ["64", "32", "48"])
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


def test_suite_copy_carries_the_reference_scripts_unmodified():
    """build/integration/sweep: run_simulations.sh and every reference source byte for byte; only main.cpp and the
    Makefile are patched copies (checked where the reference is present: the build container)."""
    suite = os.path.join(SWEEP, "nbody-sim-new")
    if not os.path.isdir(suite) or not os.path.isdir(REF):
        pytest.skip("needs /root/reference and a built build/integration/sweep")
    for name in os.listdir(REF):
        if name.endswith((".h", ".cpp", ".sh")) and name != "main.cpp" and os.path.exists(os.path.join(suite, name)):
            assert open(os.path.join(suite, name), "rb").read() == open(os.path.join(REF, name), "rb").read(), name
    assert os.path.exists(os.path.join(suite, "run_simulations.sh"))
    mk = open(os.path.join(suite, "Makefile")).read()
    ref_mk = open(os.path.join(REF, "Makefile")).read()
    # every line of the reference Makefile survives, possibly extended
    for line in ref_mk.splitlines():
        key = line.split("=")[0] if "=" in line and not line.startswith("\t") else line
        assert key in mk, line
    assert "-lnb200_methods -lnb200" in mk and "nbody-simulation-parallel_b200/csrc" in mk and "fmm_stubs.cpp" in mk


@pytest.mark.gpu
def test_reference_make_and_sweep_script(tmp_path):
    """The reference's own build + sweep: `make clean && make` with the patched Makefile, then run_simulations.sh
    (run_simulations.sh:26-60) -- here with its size list cut to two small sizes, nothing else touched -- and
    NBODY_SIM_METHODS=c so that only the CUDA method runs.  Every run must leave a BruteForce_CUDA row; the -a 1
    runs must report 100 % against the reference's own CPU brute force."""
    import shutil
    if not os.path.isdir(os.path.join(SWEEP, "nbody-sim-new")):
        pytest.skip("build/integration/sweep not built (needs /root/reference at build time)")
    work = os.path.join(tmp_path, "sweep")
    shutil.copytree(SWEEP, work)
    suite = os.path.join(work, "nbody-sim-new")
    script = open(os.path.join(suite, "run_simulations.sh")).read()
    sizes_line = 'declare -a N_VALUES=(1000 10000 100000 200000 500000 1000000 2000000 5000000)'
    assert script.count(sizes_line) == 1
    open(os.path.join(suite, "run_simulations.sh"), "w").write(script.replace(sizes_line, 'declare -a N_VALUES=(1000 4000)'))
    env = dict(os.environ, NBODY_SIM_METHODS="c", NB200=ROOT, NB200_PRECISION="64")
    r = subprocess.run(["bash", "run_simulations.sh"], cwd=suite, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "Build completed." in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "Simulation failed" not in r.stdout
    rows = []
    for f in sorted(glob.glob(os.path.join(suite, "results", "*.csv"))):
        rows += [row for row in csv.DictReader(open(f))]
    # 2 sizes x 2 dimensions without accuracy + the same with accuracy (run_simulations.sh:39-60)
    assert len(rows) == 8 and all(row["Method"] == "BruteForce_CUDA" for row in rows), rows
    assert sorted({(row["Bodies"], row["Dimension"]) for row in rows}) == [("1000", "2"), ("1000", "3"), ("4000", "2"), ("4000", "3")]
    with_acc = [row for row in rows if row.get("Accuracy(%)")]
    assert len(with_acc) == 4 and all(float(row["Accuracy(%)"]) == 100.0 for row in with_acc)


@pytest.mark.gpu
def test_steps_and_dt_flags_select_the_fused_step(tmp_path, dim=3):
    """-s / -t (added by the patch; the reference has no time loop): BruteForce_CUDA_Steps = brute_force_cuda_simulate."""
    if not os.path.exists(NBODY_SIM):
        pytest.skip("build/integration/nbody_sim not built")
    r = subprocess.run([NBODY_SIM, "-N", "3000", "-d", str(dim), "-m", "c", "-s", "25", "-t", "1e-4"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [row for f in glob.glob(os.path.join(tmp_path, "results", "*.csv")) for row in csv.DictReader(open(f))]
    assert [row["Method"] for row in rows] == ["BruteForce_CUDA", "BruteForce_CUDA_Steps"]
    assert float(rows[1]["Time(s)"]) > 0 and "25 steps of dt=0.0001" in r.stdout


def test_sweep_rows_aggregate_into_the_notebook_layout(tmp_path):
    """integration/aggregate_rows.py: rows of the patched suite's CSVs -> `Bodies,Method,Dimension,Average Runtime (s)`
    (the reference's analysis/aggregated_results.csv), duplicates of one (N, D) averaged."""
    src = tmp_path / "rows.csv"
    src.write_text("Method,Bodies,Dimension,Time(s),Accuracy(%),Precision\n"
                   "BruteForce_CUDA,1000,2,0.0002,,64\nBruteForce_CUDA,1000,2,0.0004,100.00,64\n"
                   "BruteForce_CUDA,1000,3,0.0005,,64\nBruteForce_CUDA,10000,2,0.003,,64\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "aggregate_rows.py"), str(src)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    rows = list(csv.reader(r.stdout.splitlines()))
    assert rows[0] == ["Bodies", "Method", "Dimension", "Average Runtime (s)"]
    assert rows[1][:3] == ["1000", "BruteForce_CUDA", "2"] and abs(float(rows[1][3]) - 0.0003) < 1e-12
    assert rows[2][:3] == ["10000", "BruteForce_CUDA", "2"] and rows[3][:3] == ["1000", "BruteForce_CUDA", "3"]
