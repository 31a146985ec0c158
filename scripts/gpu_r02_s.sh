#!/bin/bash
# round 2, GPU call S (1 GPU): cluster PSI kernel for large events + event-stage cuts by path-enumeration cost: full GPU suite, large-event timing, default bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/s_build.log 2>&1
timeout 170 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_events or psi_batch" > gpurun_out/s_pytest_psi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest_psi.log
timeout 120 python scripts/psi_large_timing.py > gpurun_out/s_psi_timing.txt 2>&1; echo "rc=$?" >> gpurun_out/s_psi_timing.txt
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?" >> gpurun_out/s_bench.err
