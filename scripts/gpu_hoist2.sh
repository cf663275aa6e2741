#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "hoisted or grid or full_size" 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['clocks'])); print('  ', d['roofline']['kernel'], '%.2f ms' % d['roofline']['ms_per_step'], '|', d['roofline_other'].get('kernel'), '%.2f ms' % d['roofline_other'].get('ms_per_step', 0))
except Exception as e: print('ERR', l[:400])
"; }
run "hoist serial"              LIST_B200_OVERLAP=0 $B
run "hoist overlap"             $B
run "hoist overlap 1M"          $B --chunk 1048576
run "hoist overlap v3 smem48"   LIST_B200_MLP_VARIANT=3 LIST_B200_REST_SMEM_KB=48 $B
run "hoist overlap v3 smem48 1M" LIST_B200_MLP_VARIANT=3 LIST_B200_REST_SMEM_KB=48 $B --chunk 1048576
run "hoist serial smem48"       LIST_B200_OVERLAP=0 LIST_B200_REST_SMEM_KB=48 $B
