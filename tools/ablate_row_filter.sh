#!/bin/bash
# Timing decomposition of the level-1 row filter with the DSTR_ABLATION build (results are wrong by
# construction; only the stage times are of interest).  make -C aind_smartspim_destripe_b200/csrc ablate
export DSTR_LIBRARY=$PWD/aind_smartspim_destripe_b200/lib/libdstr_ablate.so
for bits in ${ABLATE_SET:-0 1 2 4 6 8 16 24 30 31 63 95 127}; do
  DSTR_ABLATE=$bits python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('ablate',sys.argv[1],'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3))" $bits
done
