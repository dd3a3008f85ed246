#!/bin/bash
# design cost-model sweep: DSTR_NOTCH_COST_J (per mode) / DSTR_NOTCH_COST_JPAD (per 32-mode group)
for cfg in ${COST_SET:-"3.0 2.0" "3.0 0" "2.0 0" "1.0 0" "4.5 0"}; do set -- $cfg; DSTR_NOTCH_COST_J=$1 DSTR_NOTCH_COST_JPAD=$2 python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('cost',sys.argv[1],sys.argv[2],round(d['value']),'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3))" $1 $2; done
