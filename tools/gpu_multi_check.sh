#!/bin/bash
# usage: bash tools/gpu_multi_check.sh <ngpus> <outdir>   (run through gpurun --gpus <ngpus>)
G=${1:-2}; OUT=${2:-gpurun_out/multi$G}; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
python tools/e2e_breakdown.py > $OUT/e2e_breakdown.log 2>&1
python tools/e2e_breakdown.py --pageable >> $OUT/e2e_breakdown.log 2>&1
python tools/config_run.py --config c2,c3 > $OUT/configs_1gpu.jsonl 2> $OUT/configs_1gpu.err
python tools/config_run.py --config c4 --ngpus 1 >> $OUT/configs_1gpu.jsonl 2>> $OUT/configs_1gpu.err
python tools/config_run.py --config c4 --ngpus $G > $OUT/config_c4_${G}gpu.jsonl 2> $OUT/config_c4_${G}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 5 --warmup 3 > $OUT/bench_${G}gpu.json 2> $OUT/bench_${G}gpu.err
tail -3 $OUT/pytest_gpu.log; cat $OUT/e2e_breakdown.log; cut -c1-600 $OUT/configs_1gpu.jsonl; cut -c1-600 $OUT/config_c4_${G}gpu.jsonl; cat $OUT/bench_${G}gpu.json; tail -5 $OUT/*.err
