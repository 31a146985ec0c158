#!/bin/bash
# round 2, GPU call R (1 GPU): validation of the large-event PSI path and the piece-level event partition (full GPU suite, configs 2 and 3 with their parity checks)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?" >> gpurun_out/r_bench.err
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r_bench_c3.json 2> gpurun_out/r_bench_c3.err; echo "bench rc=$?" >> gpurun_out/r_bench_c3.err
