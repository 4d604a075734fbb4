#!/bin/bash
# A/B of the e2e leg's host pipeline on ONE box (PCIe rates differ between hosts):
# row chunks x steps in flight x draws queued ahead.
for cfg in "1 1 0" "2 1 2" "2 2 2" "4 2 2" "4 1 2" "1 2 2"; do
  set -- $cfg
  MB_E2E_CHUNKS=$1 MB_E2E_DEPTH=$2 MB_DRAW_AHEAD=$3 timeout 250 python bench.py --steps 5 --warmup 3 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('chunks $1 depth $2 ahead $3: e2e %.1f M/s %.2f ms' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step']))"
done
