#!/bin/bash
# same-box A/B of environment knobs (pipelined grid evaluation)
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f' % (d['value']/1e6, d['ms_per_step']), ' | '.join('%s %.2f' % (r['kernel'][:12], r['ms_per_step']) for r in [d['roofline']]+d['roofline_other']))
except Exception as e: print('ERR', l[:400])
"; }
run "default"                   $B
run "v3"                        LIST_B200_MLP_VARIANT=3 $B
run "v3 rest48"                 LIST_B200_MLP_VARIANT=3 LIST_B200_REST_SMEM_KB=48 $B
run "v3 rest24"                 LIST_B200_MLP_VARIANT=3 LIST_B200_REST_SMEM_KB=24 $B
run "v2 rest24"                 LIST_B200_REST_SMEM_KB=24 $B
run "default chunk 131072"      $B --chunk 131072
run "default chunk 524288"      $B --chunk 524288
run "default"                   $B
