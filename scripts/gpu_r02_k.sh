#!/bin/bash
# round 2, GPU call K (1 GPU): final single-GPU lines (configs 2, 1, 5, reference arm), launch list and ncu --set full captures for profiles/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/k_build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/k_smi.txt 2>&1
WQ_PROFILE=2 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?" >> gpurun_out/k_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err; echo "rc=$?" >> gpurun_out/k_bench_ref.err
timeout 600 python bench.py --config 1 --steps 5 --warmup 3 > gpurun_out/k_bench_c1.json 2> gpurun_out/k_bench_c1.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/k_bench_c5.json 2> gpurun_out/k_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/k_launches.csv \
    python bench.py --steps 2 --warmup 1 --reads 2000000 --no-parity --no-fastq > gpurun_out/k_ncu1.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:'wq_k_extend$|wq_k_seed$|wq_k_resolve$|wq_k_em_tpm_smem|wq_k_em_psi4|wq_k_em_psi$' -c 8 -f -o gpurun_out/k_full \
    python bench.py --steps 1 --warmup 0 --reads 2000000 --no-parity --no-fastq > gpurun_out/k_ncu2.log 2>&1
ls -la gpurun_out | grep " k_"
