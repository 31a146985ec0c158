#!/bin/bash
# round 2, GPU call D (2 GPUs): wq_counts_sync / wq_classes_sync under torchrun with the in-run parity check; single-GPU config-2 line with e2e_fastq
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 600 python -m pytest tests/test_fastq_ingest.py -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-parity --no-fastq > gpurun_out/d_bench_psi.json 2> gpurun_out/d_bench_psi.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --reads 5000000 --cpu-sample 600000 > gpurun_out/d_bench_2g.json 2> gpurun_out/d_bench_2g.err; echo "rc=$?" >> gpurun_out/d_bench_2g.err
WQ_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 4 --steps 3 --warmup 2 --reads 10000000 --cpu-sample 400000 > gpurun_out/d_bench_2g_c4.json 2> gpurun_out/d_bench_2g_c4.err; echo "rc=$?" >> gpurun_out/d_bench_2g_c4.err
