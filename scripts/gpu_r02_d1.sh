#!/bin/bash
# round 2, GPU call D1 (1 GPU): FASTQ ingest tests, PSI class timing, config-2 line with e2e_fastq, launch list + ncu of extend/em
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/d_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-parity --no-fastq > gpurun_out/d_bench_psi.json 2> gpurun_out/d_bench_psi.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 1 --no-parity --no-fastq > gpurun_out/d_ncu1.log 2>&1
