#!/bin/bash
# round 2, GPU call T (1 GPU): kernel times and per-phase clocks of the cluster PSI kernel; the quant stage of one rank of an 8-GPU job emulated from saved counts
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/t_build.log 2>&1
timeout 100 python scripts/psi_large_timing.py 6 12 > gpurun_out/t_psi_timing.txt 2>&1; echo "rc=$?" >> gpurun_out/t_psi_timing.txt
WQ_PROFILE=1 WQ_PSI_CLASS_TIMING=1 timeout 150 python scripts/event_stage_depth80.py 8 > gpurun_out/t_event_stage.txt 2> gpurun_out/t_event_stage.err; echo "rc=$?" >> gpurun_out/t_event_stage.txt
