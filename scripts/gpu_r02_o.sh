#!/bin/bash
# round 2, GPU call O (1 GPU): full GPU suite + config-2 / config-3 lines on the build with the cached event work list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/o_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
WQ_PROFILE=2 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?" >> gpurun_out/o_bench.err
timeout 900 python bench.py --config 3 --reads 4000000 --steps 3 --warmup 3 > gpurun_out/o_bench_c3_4m.json 2> gpurun_out/o_bench_c3_4m.err; echo "bench rc=$?" >> gpurun_out/o_bench_c3_4m.err
