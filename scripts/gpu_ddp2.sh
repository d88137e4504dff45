#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
for g in 1 0; do
  MMA_DDP_GRAPH=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$g bench.py --gpus 2 --steps 10 --warmup 3 --no-decode > gpurun_out/bench_ddp2_graph$g.json 2> gpurun_out/bench_ddp2_graph$g.err
  echo "ddp2 graph=$g -> $?"; cat gpurun_out/bench_ddp2_graph$g.json; tail -5 gpurun_out/bench_ddp2_graph$g.err
done
