for pf in 0 1184 2368 592; do
DSTR_FILTER_PREFETCH=$pf python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('prefetch',sys.argv[1],round(d['value']),'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3))" $pf
done
export DSTR_LIBRARY=$PWD/aind_smartspim_destripe_b200/lib/libdstr_ablate.so
for pf in 0 1184; do
DSTR_ABLATE=127 DSTR_FILTER_PREFETCH=$pf python bench.py --no-cpu-baseline --no-e2e --steps 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());s=d['roofline']['stage_ms_per_step'];print('ablate127 prefetch',sys.argv[1],'row',round(s['row_filter'],3),'L1',round(s['row_filter_level1'],3))" $pf
done
