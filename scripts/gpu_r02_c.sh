#!/bin/bash
# round 2, GPU call C: parity tests after the PSI / scheduling changes, config-2 bench with timeline, sanitizer logs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/c_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
WQ_PROFILE=2 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/c_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/c_racecheck.log
timeout 600 python bench.py --config 5 --reads 300000 --steps 2 --warmup 1 --cpu-sample 400000 > gpurun_out/c_bench_c5.json 2> gpurun_out/c_bench_c5.err; echo "rc=$?" >> gpurun_out/c_bench_c5.err
