for lib in libdstr_b200 libdstr_minb6 libdstr_minb8; do
DSTR_LIBRARY=$PWD/aind_smartspim_destripe_b200/lib/$lib.so python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print(sys.argv[1],round(d['value']),'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3))" $lib
done
