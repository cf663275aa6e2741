#!/bin/bash
# First GPU run of the hoisted-fc_0 path: parity tests, then A/B timings.
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "hoisted or grid or full_size" 2>&1 | tail -25 | tee gpurun_out/tests_hoist.log
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['clocks'])); print('  ', d['roofline']['kernel'], '%.2f ms' % d['roofline']['ms_per_step'], '|', d['roofline_other'].get('kernel'), '%.2f ms' % d['roofline_other'].get('ms_per_step', 0))
except Exception as e: print('ERR', l[:400])
"; }
run "plain serial"          LIST_B200_HOIST=0 LIST_B200_OVERLAP=0 $B
run "hoist serial vec4"     LIST_B200_OVERLAP=0 $B
run "hoist serial vec8"     LIST_B200_OVERLAP=0 LIST_B200_HOIST_VEC=8 $B
run "hoist overlap vec4"    $B
run "hoist overlap vec8"    LIST_B200_HOIST_VEC=8 $B
run "hoist overlap vec4 1M" $B --chunk 1048576
