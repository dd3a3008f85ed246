#!/bin/bash
# one ncu --set full capture of the level-1 mma.sync row filter launch (run under gpurun after a plain run)
TAG=${1:-r2g}
python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:filter_rows_mma -s 21 -c 1 \
    -o gpurun_out/${TAG}_rows_mma_L1 -f python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
