#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_mlp_tc.py -q -x 2>&1 | tail -3
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['clocks']))
    for r in [d['roofline']]+d['roofline_other']: print('   ', r['kernel'], '%.2f ms' % r['ms_per_step'], '%.1f %s' % (r['achieved'], r['unit']))
except Exception as e: print('ERR', l[:400])
"; }
run "hoist serial"      LIST_B200_OVERLAP=0 $B
run "hoist overlap"     $B
run "hoist overlap 1M"  $B --chunk 1048576
run "plain serial (K=3648)" LIST_B200_HOIST=0 LIST_B200_OVERLAP=0 $B
