#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): the default bench line, the per-launch list with DRAM
# bytes of a short run, and one full capture of the level-1 row filter.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r2}
python bench.py > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err || exit 1
python bench.py --workload c3 --no-extra > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err || exit 1
python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 \
    --csv --log-file gpurun_out/${TAG}_launches_bench_steps2_c2.csv python bench.py --steps 2 --warmup 1 --no-extra \
    > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:filter_rows_mma -s 21 -c 1 \
    -o gpurun_out/${TAG}_rows_mma_L1 -f python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_list.log gpurun_out/ncu_full.log
cat gpurun_out/${TAG}_bench_c2.json
