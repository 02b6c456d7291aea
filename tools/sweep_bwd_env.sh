#!/bin/bash
# sweep the hashed-replica tunables of the scatter plan; prints per-kernel times of hash_bwd / proposal_bwd
for cfg in "1 0" "2 2" "2 4" "2 6" "4 2" "4 4" "4 6" "8 2"; do
  set -- $cfg
  NRB_BWD_HASHED_COPIES=$1 NRB_BWD_HASHED_LEVELS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('copies $1 levels $2 step %.3f  hash_bwd %.3f  prop_bwd %.3f prop_fwd %.3f' % (d['ms_per_step'], k['nrb_hash_bwd:L16F2T19']['mean_ms'], k['nrb_proposal_bwd']['mean_ms'], k['nrb_proposal_fwd']['mean_ms']))"
done
