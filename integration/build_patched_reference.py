#!/usr/bin/env python
"""Build the reference's own benchmark suite with BruteForce_CUDA added as a selectable method.

Proof of the drop-in claim.  COPIES of the reference's main.cpp and Makefile
are patched at build time (never committed, never edited in place; build/ is git-ignored but travels to
the GPU box) so that

  * `#include "methods_cuda.h"` follows `#include "methods.h"`                       (main.cpp:14)
  * method letter `c` selects the CUDA brute force; with no `-m` it is in the default set, so
    run_simulations.sh (which never passes -m, run_simulations.sh:16) picks it up     (main.cpp:24-35)
  * the letter passes the validation list                                             (main.cpp:909-915)
  * a `BruteForce_CUDA` block, written in the same shape as its five CPU peers
    (main.cpp:131-183), times the call with safely_execute, computes the -a 1 accuracy
    column with compute_accuracy_omp, writes the CSV row and prints the validation forces
  * `-s <steps> -t <dt>` (new: the reference has no time loop) add a
    `BruteForce_CUDA_Steps` row: brute_force_cuda_simulate<D>(bodies, dt, steps), i.e. force +
    update_body_velocities + update_body_positions (methods.cpp:426-450) fused on the device
  * with no `-m`, the environment variable NBODY_SIM_METHODS supplies the method letters, so the
    UNMODIFIED run_simulations.sh can be pointed at one method (a full default sweep runs every CPU
    method up to N = 5e6: hours)
  * the Makefile (Makefile:1-12) gains the include paths, the link line, the one define and the stub TU
    HEAD needs to build at all, and rules that build libnb200.so (nvcc, sm_100a) and the
    adapter when they are missing.

Outputs:
  build/integration/nbody_sim                      the binary tests/test_integration.py drives
  build/integration/sweep/nbody-sim-new/           a self-contained copy of the suite (patched main.cpp + Makefile, the
                                                   reference sources and run_simulations.sh byte for byte) in which the
                                                   reference's own `make clean && make` + sweep script run
  build/integration/sweep/parlaylib/include/       the vendored headers its -I ../parlaylib/include expects
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

        // -s <steps> -t <dt>: the time-stepped run the reference's update_body_* helpers were written for
        // (methods.cpp:426-450; never called by the reference itself), fused on the device
        if (g_cuda_steps > 0) {
            std::vector<Body<D>> stepped = bodies;
            auto time_steps = safely_execute(log_output, "BruteForce_CUDA_Steps", [&]() {
                brute_force_cuda_simulate<D>(stepped, g_cuda_dt, g_cuda_steps);
                return 0;
            });
            if (time_steps >= 0) {
                csv_output << "BruteForce_CUDA_Steps," << n << "," << D << "," << std::fixed << std::setprecision(6)
                           << time_steps / 1e6;
                if (calculate_accuracy) csv_output << ",";
                csv_output << std::endl;
                log_output << g_cuda_steps << " steps of dt=" << g_cuda_dt << ": " << time_steps / 1e6 << " s (kernels: "
                           << brute_force_cuda_last_kernel_ms() / 1e3 << " s); body 0 now at";
                std::cout << g_cuda_steps << " steps of dt=" << g_cuda_dt << ": " << time_steps / 1e6 << " s (kernels: "
                          << brute_force_cuda_last_kernel_ms() / 1e3 << " s); body 0 now at";
                for (int d = 0; d < D; ++d) {
                    log_output << " " << std::setprecision(17) << stepped[0].position[d];
                    std::cout << " " << std::setprecision(17) << stepped[0].position[d];
                }
                log_output << std::endl;
                std::cout << std::endl;
            }
        }
        log_output << std::endl;
        std::cout << std::endl;
    }
'''

MAIN_EDITS = [
    # (anchor that must occur exactly once, replacement)
    ('#include "methods.h"\n',
     '#include "methods.h"\n#include "methods_cuda.h"\n#include <cstdlib>\n\n'
     '// -s / -t of the BruteForce_CUDA block (0 steps = force evaluation only, like every other method)\n'
     'static int g_cuda_steps = 0;\nstatic double g_cuda_dt = 1e-3;\n'),
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
    ('        } else if (arg == "-h" || arg == "--help") {\n',
     '        } else if ((arg == "-s" || arg == "--steps") && i + 1 < argc) {\n'
     '            g_cuda_steps = std::stoi(argv[++i]);\n'
     '        } else if ((arg == "-t" || arg == "--dt") && i + 1 < argc) {\n'
     '            g_cuda_dt = std::stod(argv[++i]);\n'
     '        } else if (arg == "-h" || arg == "--help") {\n'),
    ('            std::cout << "  -h, --help          Display this help message" << std::endl;\n',
     '            std::cout << "  -s, --steps <num>   BruteForce_CUDA: also advance the bodies <num> time steps (default: 0)" << std::endl;\n'
     '            std::cout << "  -t, --dt <dt>       time step of -s (default: 1e-3)" << std::endl;\n'
     '            std::cout << "  -h, --help          Display this help message" << std::endl;\n'),
    ("    // Get run ID from current date and time\n",
     "    // no -m on the command line: NBODY_SIM_METHODS picks the methods (lets the unmodified sweep script run one method)\n"
     "    if (methods.empty()) {\n"
     "        if (const char* env_methods = std::getenv(\"NBODY_SIM_METHODS\")) methods = env_methods;\n"
     "    }\n\n"
     "    // Get run ID from current date and time\n"),
    ("    return 0;\n}\n", "    brute_force_cuda_release();   // tear the device context down before the CUDA runtime's own atexit\n    return 0;\n}\n"),
]

MAKEFILE_EDITS = [
    ("CXX = g++\n",
     "# BruteForce_CUDA: the B200-native brute-force path (libnb200.so + its C++ adapter), built on demand\n"
     "NB200 ?= $(abspath ../../../..)\n"
     "NB200_LIB = $(NB200)/nbody-simulation-parallel_b200/lib\n"
     "CXX = g++\n"),
    ("CXXFLAGS = -std=c++17 -O3 -fopenmp -I ../parlaylib/include\n",
     "CXXFLAGS = -std=c++17 -O3 -fopenmp -I ../parlaylib/include -I. -I$(NB200)/include -I$(NB200)/nbody-simulation-parallel_b200/host \\\n"
     "           '-DMultipoleExpansion=Expansion<D,10>' -w    # the define and fmm_stubs.cpp: HEAD does not build without them (fmm_omp.cpp:228, fmm.h:98,101)\n"),
    ("LDFLAGS = -fopenmp\n",
     "LDFLAGS = -fopenmp -L$(NB200_LIB) -lnb200_methods -lnb200 -Wl,-rpath,$(NB200_LIB)\n"),
    ("SOURCES = main.cpp methods.cpp\n", "SOURCES = main.cpp methods.cpp fmm_stubs.cpp\n"),
    ("$(EXECUTABLE): $(OBJECTS)\n",
     "# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared ... csrc/nb200_api.cu (csrc/Makefile), then the adapter\n"
     "$(NB200_LIB)/libnb200.so:\n"
     "\t$(MAKE) -C $(NB200)/nbody-simulation-parallel_b200/csrc\n"
     "$(NB200_LIB)/libnb200_methods.so: $(NB200_LIB)/libnb200.so\n"
     "\t$(MAKE) -C $(NB200)/nbody-simulation-parallel_b200/host $(NB200_LIB)/libnb200_methods.so\n\n"
     "$(EXECUTABLE): $(OBJECTS) $(NB200_LIB)/libnb200_methods.so\n"),
]


def patch(text: str, edits, what: str) -> str:
    for anchor, repl in edits:
        if text.count(anchor) != 1:
            raise SystemExit(f"{what}: patch anchor not found exactly once (reference changed?): {anchor!r}")
        text = text.replace(anchor, repl)
    return text


def main():
    if not os.path.exists(os.path.join(REF, "main.cpp")):
        print(f"reference not present at {REF}: keeping prebuilt build/integration (if any)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(REF, "main.cpp")) as f:
        patched_main = patch(f.read(), MAIN_EDITS, "main.cpp")
    src = os.path.join(OUT, "main_with_cuda.cpp")
    with open(src, "w") as f:
        f.write(patched_main)
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

    # ---- the self-contained suite copy for the reference's own build + sweep scripts
    suite = os.path.join(OUT, "sweep", "nbody-sim-new")
    if os.path.isdir(os.path.join(OUT, "sweep")):
        shutil.rmtree(os.path.join(OUT, "sweep"))
    os.makedirs(suite)
    for name in sorted(os.listdir(REF)):
        if name.endswith((".h", ".cpp")) and name != "main.cpp":
            shutil.copy(os.path.join(REF, name), os.path.join(suite, name))
    shutil.copy(os.path.join(REF, "run_simulations.sh"), os.path.join(suite, "run_simulations.sh"))   # byte for byte
    shutil.copy(os.path.join(ROOT, "integration", "fmm_stubs.cpp"), os.path.join(suite, "fmm_stubs.cpp"))
    with open(os.path.join(suite, "main.cpp"), "w") as f:
        f.write(patched_main)
    with open(os.path.join(REF, "Makefile")) as f:
        mk = patch(f.read(), MAKEFILE_EDITS, "Makefile")
    with open(os.path.join(suite, "Makefile"), "w") as f:
        f.write(mk)
    shutil.copytree(os.path.join(REF, "..", "parlaylib", "include"), os.path.join(OUT, "sweep", "parlaylib", "include"))
    print("wrote", suite)
    return 0


if __name__ == "__main__":
    sys.exit(main())
