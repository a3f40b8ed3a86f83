python -m pytest tests -m gpu -x -q -k "symmetric" 2>&1 | tail -3
for n in 262144 1048576; do python tools/run_case.py --n $n --steps 2 --warmup 1 --opt symmetric=1; done
