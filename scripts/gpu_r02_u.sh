#!/bin/bash
# round 2, GPU call U (1 GPU): final check of the round's last state: full GPU suite, default bench, large-event PSI timing (block / cluster kernels,
# lane-split shares phase on and off, per-phase clocks), the event stage of one rank of an 8-GPU job emulated from saved counts
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/u_build.log 2>&1
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
timeout 110 python bench.py --steps 5 --warmup 3 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?" >> gpurun_out/u_bench.err
timeout 40 python scripts/psi_large_timing.py 6 > gpurun_out/u_psi_timing.txt 2>&1; echo "rc=$?" >> gpurun_out/u_psi_timing.txt
WQ_PSI_CLASS_TIMING=1 timeout 60 python scripts/event_stage_depth80.py 8 > gpurun_out/u_event_stage.txt 2> gpurun_out/u_event_stage.err; echo "rc=$?" >> gpurun_out/u_event_stage.txt
