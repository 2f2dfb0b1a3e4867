#!/bin/bash
# A/B of library builds under bench.py's headline schedule: [AB_FLAGS="--two-streams"] profiles/ab_bench.sh <reps> <steps> <lib> [<lib> ...]
reps=$1; steps=$2; shift 2
F="$AB_FLAGS --steps $steps --no-cpu-baseline --no-e2e --no-extra --no-updates --no-secondary --no-small --no-parity-check"
for r in $(seq 1 $reps); do
  for lib in "$@"; do
    FDQL_LIB=$lib python bench.py $F 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; k=r['kernels']; a=r.get('alone') or k
print('$lib rep $r: ms_pass %.4f value %.4g | in the timed region:' % (j['ms_per_pass'], j['value']), ' '.join('%s %.4f' % (n, d['ms']) for n, d in k.items()),
      '| alone:', ' '.join('%s %.4f' % (n, d['ms_alone']) for n, d in a.items()))"
  done
done
