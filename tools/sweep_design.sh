#!/bin/bash
# Row-filter time against the weights of the notch design's cost model (modes of the rank-J correction vs FIR taps).
for cj in ${CJ:-3 1 2 4}; do for cp in ${CP:-0.5 0.1 1 2}; do
  DSTR_NOTCH_COST_J=$cj DSTR_NOTCH_COST_JPAD=$cp python bench.py --no-extra --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_step']; print('cost_j $cj cost_jpad $cp', round(d['ms_per_step'],3), round(s['row_filter'],3), round(s['row_filter_level1'],3))"
done; done
