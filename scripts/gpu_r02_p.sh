#!/bin/bash
# round 2, GPU call P (8 GPUs): weak scaling of config 2 and strong scaling of config 4 with the NCCL exchange inside the library
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/p_build.log 2>&1
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/p_bench_8g.json 2> gpurun_out/p_bench_8g.err; echo "rc=$?" >> gpurun_out/p_bench_8g.err
WQ_PROFILE=1 NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config 4 --steps 3 --warmup 3 > gpurun_out/p_bench_8g_c4.json 2> gpurun_out/p_bench_8g_c4.err; echo "rc=$?" >> gpurun_out/p_bench_8g_c4.err
nproc > gpurun_out/p_host.txt
