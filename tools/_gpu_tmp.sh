for n in 262144 1048576; do for ti in 4 8; do python tools/run_case.py --n $n --steps 2 --warmup 1 --opt symmetric=1 --opt sym_ti=$ti; done; done
python tools/run_case.py --n 131072 --dim 2 --steps 2 --warmup 1 --opt symmetric=1 --opt sym_ti=8
