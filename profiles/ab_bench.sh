#!/bin/bash
# A/B of library builds under bench.py's headline schedule: profiles/ab_bench.sh <reps> <steps> <lib> [<lib> ...]
reps=$1; steps=$2; shift 2
F="$AB_FLAGS --steps $steps --no-cpu-baseline --no-e2e --no-extra --no-updates --no-secondary --no-small --no-parity-check"
for r in $(seq 1 $reps); do
  for lib in "$@"; do
    FDQL_LIB=$lib python bench.py $F 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernels']
print('$lib rep $r: ms_pass %.4f value %.4g gather %.4f/%.4f tqc %.4f/%.4f' % (j['ms_per_pass'], j['value'], k['sample_gather_kernel']['ms'], k['sample_gather_kernel']['ms_alone'], k['tqc_loss_kernel']['ms'], k['tqc_loss_kernel']['ms_alone']))"
  done
done
