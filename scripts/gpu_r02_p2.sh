#!/bin/bash
# round 2, GPU call P2 (8 GPUs): weak scaling of config 2 after the piece-level event partition and the large-event PSI path
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p2_build.log 2>&1
WQ_PROFILE=1 NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/p2_bench_8g.json 2> gpurun_out/p2_bench_8g.err; echo "rc=$?" >> gpurun_out/p2_bench_8g.err
