#!/bin/bash
# Multi-GPU bench lines on one box (run under `gpurun --gpus N`): c2 (resident + end-to-end, all-rank PCIe probe),
# c4 (one streamed tile split into Z-slabs) and, at N = 8, c5 (one streamed tile per GPU).  Lines land in gpurun_out/.
N=${1:-2}; TAG=${2:-r2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_${N}gpu.json 2> gpurun_out/${TAG}_bench_c2_${N}gpu.err; echo "c2 rc=$?"
run 29512 --workload c4 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c4_${N}gpu.json 2> gpurun_out/${TAG}_bench_c4_${N}gpu.err; echo "c4 rc=$?"
if [ "$N" = "8" ]; then
  run 29513 --workload c5 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c5_${N}gpu.json 2> gpurun_out/${TAG}_bench_c5_${N}gpu.err; echo "c5 rc=$?"
fi
nproc; free -g | head -2
for f in gpurun_out/${TAG}_bench_c*_${N}gpu.json; do echo "== $f"; cut -c1-600 $f; done
