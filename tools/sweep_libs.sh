#!/bin/bash
# bench the row filter with every build variant found in aind_smartspim_destripe_b200/lib
for lib in aind_smartspim_destripe_b200/lib/libdstr_b200.so aind_smartspim_destripe_b200/lib/libdstr_var_*.so; do
DSTR_LIBRARY=$PWD/$lib python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print(sys.argv[1].split('/')[-1],round(d['value']),'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3),'hist',round(s['histogram'],3),'an',round(s['analysis_l1'],3),'fin',round(s['final_synthesis_epilogue'],3))" $lib
done
