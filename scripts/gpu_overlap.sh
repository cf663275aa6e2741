#!/bin/bash
# A/B of the two-stream chunk pipeline (api.cu run_chunks): serial vs overlapped, MLP ring depth, chunk size.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
echo "== parity (grid + shards)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['clocks']))
except Exception as e: print('ERR', l[:300])
"; }
run "serial v2 chunk 262144"   LIST_B200_OVERLAP=0 $B
run "overlap v2 chunk 262144"  LIST_B200_OVERLAP=1 $B
run "overlap v3 chunk 262144"  LIST_B200_OVERLAP=1 LIST_B200_MLP_VARIANT=3 $B
run "serial v3 chunk 262144"   LIST_B200_OVERLAP=0 LIST_B200_MLP_VARIANT=3 $B
run "overlap v3 chunk 1048576" LIST_B200_OVERLAP=1 LIST_B200_MLP_VARIANT=3 $B --chunk 1048576
run "overlap v2 chunk 1048576" LIST_B200_OVERLAP=1 $B --chunk 1048576
run "overlap v3 chunk 524288"  LIST_B200_OVERLAP=1 LIST_B200_MLP_VARIANT=3 $B --chunk 524288
