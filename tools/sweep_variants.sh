#!/bin/bash
# Runs bench.py against every experimental build under build/variants (launch-bounds sweep).
for lib in build/variants/*.so; do
  RLS_B200_LIB=$PWD/$lib python bench.py --no-cpu --steps 10 --warmup 3 --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
o=d.get('other_workloads',{})
print('$lib', 'diel %.2f G/s' % (d['value']/1e9), ' '.join('%s %.2f' % (k.split('_')[0]+k.split('_')[1][:3], v['samples_per_s']/1e9) for k,v in o.items()))
"
done
