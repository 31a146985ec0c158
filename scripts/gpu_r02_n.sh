#!/bin/bash
# round 2, GPU call N (1 GPU): 4-bit quality dictionary batches (GPU test, e2e leg), full GPU suite, config-2 line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?" >> gpurun_out/n_bench.err
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 > gpurun_out/n_bench_c4.json 2> gpurun_out/n_bench_c4.err; echo "bench rc=$?" >> gpurun_out/n_bench_c4.err
