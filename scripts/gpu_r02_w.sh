#!/bin/bash
# round 2, GPU call W (1 GPU, the round's last GPU seconds): cluster PSI kernel with the bit-mask prop_sum phase: the GPU tests that reach it, kernel times and per-phase clocks
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/w_build.log 2>&1
timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_gpu_branches.py -m gpu -q -k "psi or beyond_1024" > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
timeout 25 python scripts/psi_large_timing.py 6 > gpurun_out/w_psi_timing.txt 2>&1; echo "rc=$?" >> gpurun_out/w_psi_timing.txt
