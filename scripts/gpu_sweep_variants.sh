#!/usr/bin/env bash
# Developer experiment: run scripts/gpu_sweep_rollout.py for the product library and for every variant build named on the
# command line (openkite_b200/_variants/<tag>).  Usage: bash scripts/gpu_sweep_variants.sh <out.log> <cases...> -- <tags...>
OUT=$1; shift
CASES=(); while [ "$1" != "--" ] && [ $# -gt 0 ]; do CASES+=("$1"); shift; done; shift
echo "== product" >> $OUT; python scripts/gpu_sweep_rollout.py "${CASES[@]}" 2>&1 | grep -E "^B=|peak" >> $OUT
for t in "$@"; do echo "== variant $t" >> $OUT; KITE_VARIANT=$t python scripts/gpu_sweep_rollout.py "${CASES[@]}" 2>&1 | grep -E "^B=|peak" >> $OUT; done
cat $OUT
