#!/usr/bin/env python
"""Build the reference's own benchmark driver with BruteForce_CUDA added as a selectable method.

Proof of the drop-in claim: a COPY of the reference's main.cpp is
patched at build time (never committed, never edited in place) so that

  * `#include "methods_cuda.h"` follows `#include "methods.h"`                     (main.cpp:14)
  * method letter `c` selects the CUDA brute force; with no `-m` it is in the default set, so
    run_simulations.sh (which never passes -m, run_simulations.sh:16) picks it up    (main.cpp:24-35)
  * the letter passes the validation list                                           (main.cpp:909-915)
  * a `BruteForce_CUDA` block, written in the same shape as its five CPU peers
    (main.cpp:131-183), times the call with safely_execute, computes the -a 1 accuracy
    column with compute_accuracy_omp, writes the CSV row and prints the validation forces.

The CSV label `BruteForce_CUDA` is the one the reference's notebook already uses for its
hand-pasted rows (analysis/aggregated_results.csv:227-234).

Output: build/integration/nbody_sim (+ the patched source beside it; build/ is git-ignored but
travels to the GPU box).  Flags are the reference Makefile's (Makefile:2-3) plus the one define
HEAD needs to compile and the stub objects it needs to link.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference/nbody-sim-new")
OUT = os.path.join(ROOT, "build", "integration")
PKG = os.path.join(ROOT, "nbody-simulation-parallel_b200")

CUDA_BLOCK = r'''
    // Brute force on the GPU(s): B200-native libnb200 behind the methods.h-style entry point.
    // Selected with -m c, and part of the default method set (also above the 1M-body limit the
    // CPU brute-force variants are skipped at).
    if (run_cuda) {
        log_output << "Brute force O(n²) CUDA (B200, libnb200) approach:" << std::endl;
        std::cout << "Brute force O(n²) CUDA (B200, libnb200) approach:" << std::endl;

        try { brute_force_cuda_warmup<D>(bodies.size()); } catch (const std::exception&) {}
        std::vector<Vector<D>> forces_bf_cuda;
        auto time_bf_cuda = safely_execute(log_output, "BruteForce_CUDA", [&]() {
            forces_bf_cuda = brute_force_cuda_n_body<D>(bodies);
            return forces_bf_cuda;
        });

        double accuracy_bf_cuda = -1.0;
        if (calculate_accuracy && time_bf_cuda >= 0) {
            accuracy_bf_cuda = compute_accuracy_omp(forces_bf_cuda, reference_forces);
        }
        double time_bf_cuda_seconds = time_bf_cuda / 1e6;

        if (time_bf_cuda >= 0) {
            csv_output << "BruteForce_CUDA," << n << "," << D;
            if (time_bf_cuda_seconds < 1e-6) {
                csv_output << "," << std::scientific << std::setprecision(6) << time_bf_cuda_seconds;
            } else {
                csv_output << "," << std::fixed << std::setprecision(6) << time_bf_cuda_seconds;
            }
            if (calculate_accuracy) {
                csv_output << "," << std::fixed << std::setprecision(2) << accuracy_bf_cuda;
            }
            csv_output << std::endl;

            log_output << "Time taken: " << time_bf_cuda_seconds << " s (kernel only: "
                       << brute_force_cuda_last_kernel_ms() / 1e3 << " s)" << std::endl;
            std::cout << "Time taken: " << time_bf_cuda_seconds << " s (kernel only: "
                      << brute_force_cuda_last_kernel_ms() / 1e3 << " s)" << std::endl;
            if (calculate_accuracy) {
                log_output << "Accuracy: " << std::to_string(accuracy_bf_cuda) + "%" << std::endl;
                std::cout << "Accuracy: " << std::to_string(accuracy_bf_cuda) + "%" << std::endl;
            }
            if (n >= 3) {
                print_validation_forces(forces_bf_cuda, n, log_output);
                print_validation_forces(forces_bf_cuda, n, std::cout);
            }
        }
        log_output << std::endl;
        std::cout << std::endl;
    }
'''

EDITS = [
    # (anchor that must occur exactly once, replacement)
    ('#include "methods.h"\n', '#include "methods.h"\n#include "methods_cuda.h"\n'),
    ("    bool run_fmm = methods.find('f') != std::string::npos;\n",
     "    bool run_fmm = methods.find('f') != std::string::npos;\n"
     "    bool run_cuda = methods.find('c') != std::string::npos;\n"),
    ("        run_fmm = true;\n    }\n", "        run_fmm = true;\n        run_cuda = true;\n    }\n"),
    ('    if (run_bruteforce) std::cout << "Brute Force ";\n',
     '    if (run_bruteforce) std::cout << "Brute Force ";\n    if (run_cuda) std::cout << "Brute Force CUDA ";\n'),
    ("    // Barnes-Hut methods - only run if specified\n",
     CUDA_BLOCK + "\n    // Barnes-Hut methods - only run if specified\n"),
    ("                if (c != 'a' && c != 'b' && c != 'h' && c != 'f') {",
     "                if (c != 'a' && c != 'b' && c != 'h' && c != 'f' && c != 'c') {"),
    ('Valid methods: a=bruteforce, b=barnes-hut, h=bvh, f=fmm"',
     'Valid methods: a=bruteforce, c=bruteforce CUDA (B200), b=barnes-hut, h=bvh, f=fmm"'),
]


def patch(text: str) -> str:
    for anchor, repl in EDITS:
        if text.count(anchor) != 1:
            raise SystemExit(f"patch anchor not found exactly once (reference changed?): {anchor!r}")
        text = text.replace(anchor, repl)
    return text


def main():
    if not os.path.exists(os.path.join(REF, "main.cpp")):
        print(f"reference not present at {REF}: keeping prebuilt build/integration (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(OUT, "main_with_cuda.cpp")
    with open(os.path.join(REF, "main.cpp")) as f:
        patched = patch(f.read())
    with open(src, "w") as f:
        f.write(patched)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "host")], check=True)
    libdir = os.path.join(PKG, "lib")
    cxx = ["g++", "-std=c++17", "-O3", "-fopenmp", "-w", f"-I{REF}/../parlaylib/include", f"-I{REF}",
           f"-I{ROOT}/include", f"-I{PKG}/host", "-DMultipoleExpansion=Expansion<D,10>"]
    obj = os.path.join(OUT, "main_with_cuda.o")
    subprocess.run(cxx + ["-c", src, "-o", obj], check=True)
    # methods.o (the reference's own, unmodified) + the FMM stub TU come from oracle/_ref
    ref_objs = [os.path.join(ROOT, "oracle", "_ref", "methods.o"), os.path.join(ROOT, "oracle", "_ref", "ref_shim.o")]
    exe = os.path.join(OUT, "nbody_sim")
    subprocess.run(["g++", "-fopenmp", "-o", exe, obj] + ref_objs +
                   [f"-L{libdir}", "-lnb200_methods", "-lnb200", f"-Wl,-rpath,{libdir}"], check=True)
    print("built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
