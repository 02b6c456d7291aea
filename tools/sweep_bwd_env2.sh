#!/bin/bash
# sweep the dense-lattice replica scale with hashed replicas off (after run merging)
for sc in 0 0.5 1 2 4; do
  NRB_BWD_HASHED_COPIES=1 NRB_BWD_SCALE=$sc python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('scale $sc step %.3f  hash_bwd %.3f  prop_bwd %.3f launches %d' % (d['ms_per_step'], k['nrb_hash_bwd:L16F2T19']['mean_ms'], k['nrb_proposal_bwd']['mean_ms'], d['gpu_launches']))"
done
