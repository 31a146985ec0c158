#!/bin/bash
# round 2, GPU call A: parity tests, config-2 bench with the host-stage timeline, launch list, full ncu capture of the alignment kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/a_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/a_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
WQ_PROFILE=2 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 1 --warmup 1 --reads 2000000 --no-parity > gpurun_out/a_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'wq_k_extend$|wq_k_seed$|wq_k_resolve$|wq_k_extend2$' -c 4 -f -o gpurun_out/a_align \
    python bench.py --steps 1 --warmup 0 --reads 2000000 --no-parity > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out | tail -20
