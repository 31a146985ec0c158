#!/bin/bash
# round 2, GPU call G (1 GPU): native index file (GPU parity test, bench handles loaded from the file), config 1 and 3 lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/g_build.log 2>&1
timeout 900 python -m pytest tests/test_index_file.py -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?" >> gpurun_out/g_bench.err
timeout 900 python bench.py --config 3 --steps 3 --warmup 3 > gpurun_out/g_bench_c3.json 2> gpurun_out/g_bench_c3.err; echo "bench rc=$?" >> gpurun_out/g_bench_c3.err
timeout 600 python bench.py --config 1 --steps 3 --warmup 3 > gpurun_out/g_bench_c1.json 2> gpurun_out/g_bench_c1.err; echo "bench rc=$?" >> gpurun_out/g_bench_c1.err
