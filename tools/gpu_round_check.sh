#!/bin/bash
# bash tools/gpu_round_check.sh [outdir]   -- the 1-GPU evidence of a round: GPU tests, bench (both arms), the ncu
# launch list of the same bench command, one ncu --set full capture of the hot kernel
set -x
OUT=${1:-gpurun_out/round}; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-fp64 > $OUT/ncu_bench.log 2>&1
python tools/run_case.py --n 1048576 --steps 1 --warmup 2 && ncu --set full --clock-control none --import-source on -k regex:nb_force -s 2 -c 1 -o $OUT/prof_f32_sym_n1m python tools/run_case.py --n 1048576 --steps 1 --warmup 2 > $OUT/ncu_f32.log 2>&1
tail -3 $OUT/pytest_gpu.log
cat $OUT/bench.json
