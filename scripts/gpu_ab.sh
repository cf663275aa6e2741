#!/bin/bash
# same-box A/B of library builds in ab/ against the in-tree build
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
run() { echo "== $1"; shift; env "$@" 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.readline()
try:
    d=json.loads(l); print('value %.1f Mq/s  ms/step %.2f  clocks %s' % (d['value']/1e6, d['ms_per_step'], d['clocks']))
    print('   ', ' | '.join('%s %.2f ms' % (r['kernel'], r['ms_per_step']) for r in [d['roofline']]+d['roofline_other']))
except Exception as e: print('ERR', l[:400])
"; }
for v in "$@"; do
  if [ "$v" = "tree" ]; then L=""; else L="LIST_B200_LIB=$PWD/ab/liblist_$v.so"; fi
  run "$v serial"  $L LIST_B200_OVERLAP=0 $B
  run "$v overlap" $L $B
done
