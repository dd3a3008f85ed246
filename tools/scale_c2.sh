#!/bin/bash
# c2 line only (resident + end-to-end) on N GPUs of one box
N=${1:-8}; TAG=${2:-r2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --no-cpu-baseline \
  > gpurun_out/${TAG}_bench_c2_${N}gpu.json 2> gpurun_out/${TAG}_bench_c2_${N}gpu.err; echo "rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench_c2_${N}gpu.json
