#!/bin/bash
# round 2, GPU call L (1 GPU): extension-kernel variants (top-level extension out of line; occupancy), launch list at the full bench size, bias-mode cost
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/l_build.log 2>&1
WQ_BENCH_READS=4000000 timeout 900 python scripts/kernel_timing.py whippet.jl_b200/csrc/variants/base.so whippet.jl_b200/csrc/variants/nested.so whippet.jl_b200/csrc/variants/nested_mb6.so whippet.jl_b200/csrc/variants/mb6.so whippet.jl_b200/csrc/variants/mb10.so whippet.jl_b200/csrc/variants/nested_mb10.so whippet.jl_b200/csrc/variants/base.so whippet.jl_b200/csrc/variants/nested.so > gpurun_out/l_variants.log 2>&1
timeout 900 python scripts/bias_timing.py > gpurun_out/l_bias.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/l_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-parity --no-fastq > gpurun_out/l_ncu1.log 2>&1
