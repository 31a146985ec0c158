#!/bin/bash
# round 2, GPU call E (1 GPU): streaming SAM test, FASTQ ingest with small batches + pinned cache (e2e_fastq), phase timers
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
timeout 900 python -m pytest tests/test_fastq_ingest.py tests/test_sam.py -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
WQ_FASTQ_PROF=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-parity > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
nproc > gpurun_out/e_host.txt; lscpu | head -20 >> gpurun_out/e_host.txt
