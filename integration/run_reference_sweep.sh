#!/bin/bash
# Runs the reference's OWN sweep script, unmodified (run_simulations.sh: `make clean && make`, then N in
# {1e3 .. 5e6} x D in {2,3}, and the first four sizes again with -a 1), inside the suite copy that
# integration/build_patched_reference.py wrote, with NBODY_SIM_METHODS=c so that the default method set is the
# CUDA brute force only (every CPU method up to N = 5e6 would take hours).  One pass per precision.
#   integration/run_reference_sweep.sh [out_dir]          (default gpurun_out/sweep)
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-$ROOT/gpurun_out/sweep}"
SRC="$ROOT/build/integration/sweep"
[ -d "$SRC/nbody-sim-new" ] || { echo "$SRC is missing: run integration/build_patched_reference.py where /root/reference exists"; exit 1; }
mkdir -p "$OUT"
OUT="$(cd "$OUT" && pwd)"
for PREC in 64 32; do
    WORK="$(mktemp -d)"
    cp -r "$SRC/." "$WORK/"
    ( cd "$WORK/nbody-sim-new" && cmp run_simulations.sh "$SRC/nbody-sim-new/run_simulations.sh" &&
      NBODY_SIM_METHODS=c NB200="$ROOT" NB200_PRECISION=$PREC NB200_TRACE=1 bash run_simulations.sh \
          > "$OUT/run_simulations_fp$PREC.log" 2> "$OUT/run_simulations_fp$PREC.trace" )
    echo "Method,Bodies,Dimension,Time(s),Accuracy(%),Precision" > "$OUT/bruteforce_cuda_rows_fp$PREC.csv"
    for f in "$WORK"/nbody-sim-new/results/*.csv; do
        tail -n +2 "$f" | awk -F, -v p=$PREC '{ acc = (NF >= 5) ? $5 : ""; print $1","$2","$3","$4","acc","p }' >> "$OUT/bruteforce_cuda_rows_fp$PREC.csv"
    done
    rm -rf "$WORK"
done
grep -h "nb200 trace" "$OUT"/run_simulations_fp*.trace > "$OUT/phase_trace.txt" || true
echo "sweep rows in $OUT"
