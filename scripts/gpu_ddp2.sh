#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 20 --warmup 5 --no-decode > gpurun_out/bench_ddp2.json 2> gpurun_out/bench_ddp2.err
echo "ddp2 -> $?"; cat gpurun_out/bench_ddp2.json; tail -3 gpurun_out/bench_ddp2.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus ${NG:-2} --steps 2 --warmup 1 > gpurun_out/bench_ddp2_ref.json 2> gpurun_out/bench_ddp2_ref.err
echo "ddp2 reference -> $?"; cut -c1-200 gpurun_out/bench_ddp2_ref.json
