#!/bin/bash
# round 2, GPU call I (1 GPU): full GPU suite (bias parity, index file, streaming SAM), PSI class timing on configs 3 and 5, bench lines 2 / 3 / 5
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/i_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 900 python bench.py --config 3 --steps 1 --warmup 1 --no-parity > gpurun_out/i_bench_c3_psi.json 2> gpurun_out/i_bench_c3_psi.err
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-parity > gpurun_out/i_bench_c5_psi.json 2> gpurun_out/i_bench_c5_psi.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?" >> gpurun_out/i_bench.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/i_bench_c3.json 2> gpurun_out/i_bench_c3.err; echo "bench rc=$?" >> gpurun_out/i_bench_c3.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/i_bench_c5.json 2> gpurun_out/i_bench_c5.err; echo "bench rc=$?" >> gpurun_out/i_bench_c5.err
