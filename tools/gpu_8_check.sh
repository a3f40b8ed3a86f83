#!/bin/bash
# bash tools/gpu_8_check.sh   (run through gpurun --gpus 8)
OUT=gpurun_out/multi8c; mkdir -p $OUT
T=tests/test_gpu_parity.py::test_single_process_multi_gpu_matches_one_gpu
timeout 900 python -m pytest -x -q "$T[64-4]" "$T[32-4]" "$T[64-8]" "$T[32-8]" > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi.log
for G in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 5 --warmup 3 > $OUT/bench_${G}gpu.json 2> $OUT/bench_${G}gpu.err
done
timeout 300 python tools/config_run.py --config c4 --ngpus 8 > $OUT/config_c4_8gpu.jsonl 2> $OUT/config_c4_8gpu.err
timeout 300 python tools/config_run.py --config c5 --ngpus 8 > $OUT/config_c5_8gpu.jsonl 2> $OUT/config_c5_8gpu.err
tail -3 $OUT/pytest_multi.log; for f in $OUT/bench_*gpu*.json; do cut -c1-200 $f; done; cut -c1-300 $OUT/config_c*.jsonl; tail -n 3 $OUT/*.err
