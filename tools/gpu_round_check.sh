set -x
mkdir -p gpurun_out/r01d
python -m pytest tests -m gpu -x -q > gpurun_out/r01d/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01d/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r01d/bench.json 2> gpurun_out/r01d/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d/bench_ref.json 2> gpurun_out/r01d/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-fp64 > gpurun_out/r01d/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nb_force -s 2 -c 1 -o gpurun_out/r01d/prof_f32_sym_n1m python tools/run_case.py --n 1048576 --steps 1 --warmup 2 > gpurun_out/r01d/ncu_f32.log 2>&1
tail -3 gpurun_out/r01d/pytest_gpu.log
cat gpurun_out/r01d/bench.json
python tools/config_run.py --config c2,c3,c4 > gpurun_out/r01d/configs_1gpu.jsonl 2> gpurun_out/r01d/configs_1gpu.err
python integration/run_sweep.py --out gpurun_out/r01d/sweep > gpurun_out/r01d/sweep.log 2>&1
tail -3 gpurun_out/r01d/sweep.log
