#!/bin/bash
# bash tools/gpu_8_check.sh   (run through gpurun --gpus 8)
OUT=gpurun_out/multi8; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -m pytest tests -m gpu -x -q -k "multi_gpu or torchrun" > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi.log
for G in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 5 --warmup 3 > $OUT/bench_${G}gpu.json 2> $OUT/bench_${G}gpu.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 --no-fp64 --opt exchange=0 > $OUT/bench_8gpu_nccl.json 2> $OUT/bench_8gpu_nccl.err
python tools/config_run.py --config c4 --ngpus 8 > $OUT/config_c4_8gpu.jsonl 2> $OUT/config_c4_8gpu.err
python tools/config_run.py --config c4 --ngpus 4 > $OUT/config_c4_4gpu.jsonl 2> $OUT/config_c4_4gpu.err
python tools/config_run.py --config c5 --ngpus 8 > $OUT/config_c5_8gpu.jsonl 2> $OUT/config_c5_8gpu.err
tail -3 $OUT/pytest_multi.log; for f in $OUT/bench_*gpu*.json; do cut -c1-200 $f; done; cut -c1-400 $OUT/config_c*.jsonl; tail -n 5 $OUT/*.err
