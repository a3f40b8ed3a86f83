timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for n in 262144 1048576; do timeout 200 python tools/run_case.py --n $n --precision 64 --steps 1 --warmup 1; done
timeout 100 python tools/run_case.py --n 131072 --dim 2 --precision 64 --steps 2 --warmup 1
for ti in 4 8; do timeout 100 python tools/run_case.py --n 131072 --dim 2 --steps 3 --warmup 1 --opt sym_ti=$ti; done
