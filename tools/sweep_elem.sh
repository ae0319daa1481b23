#!/bin/bash
# tuning sweep of the elementwise TR kernel (unroll x CTAs/SM); prints GB/s per setting
for u in 2 4 8; do for c in 4 6 8 12 16; do
  echo "unroll=$u ctas=$c"
  TQ_ELEM_UNROLL=$u TQ_ELEM_CTAS=$c python tools/microbench.py --sizes 51380224 268435456 2>&1 | grep "k=3 b=9" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   n=%d best %.0f GB/s med %.0f' % (d['n'], d['GBs_best'], d['GBs_med']))"
done; done
