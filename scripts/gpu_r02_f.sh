#!/bin/bash
# round 2, GPU call F (1 GPU): new PSI EM kernels (block per event with pair-parallel shares; 4-lane kernel with register-resident psi)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/f_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-parity --no-fastq > gpurun_out/f_bench_psi.json 2> gpurun_out/f_bench_psi.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
timeout 900 python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/f_bench_c5.json 2> gpurun_out/f_bench_c5.err; echo "bench rc=$?" >> gpurun_out/f_bench_c5.err
