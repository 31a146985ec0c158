#!/bin/bash
# round 2, GPU call M (1 GPU): fixed_len batches (GPU test, e2e leg of the bench without offset arrays)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_branches.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?" >> gpurun_out/m_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/m_smoke.log
