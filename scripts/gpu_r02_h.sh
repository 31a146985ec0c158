#!/bin/bash
# round 2, GPU call H (2 GPUs): config-2 line from the native index file; 2-GPU weak (config 2) and strong (config 4) lines with the in-library NCCL exchange
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/h_build.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?" >> gpurun_out/h_bench.err
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/h_bench_2g.json 2> gpurun_out/h_bench_2g.err; echo "rc=$?" >> gpurun_out/h_bench_2g.err
WQ_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 4 --steps 3 --warmup 3 --reads 10000000 --cpu-sample 400000 > gpurun_out/h_bench_2g_c4.json 2> gpurun_out/h_bench_2g_c4.err; echo "rc=$?" >> gpurun_out/h_bench_2g_c4.err
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/h_bench_c3.json 2> gpurun_out/h_bench_c3.err; echo "bench rc=$?" >> gpurun_out/h_bench_c3.err
