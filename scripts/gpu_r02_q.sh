#!/bin/bash
# round 2, GPU call Q (1 GPU): block sizes of the PSI block kernel's launch classes (sweep latency of the classes, serialised by the timing switch)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/q_build.log 2>&1
for t in 32,64,128,256 64,64,128,256 64,128,256,512 128,128,256,512 32,32,64,128 128,256,512,1024; do
  WQ_PSI_THREADS=$t WQ_PROFILE=2 WQ_PSI_CLASS_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-parity --no-fastq --reads 10000000 > /dev/null 2> gpurun_out/q_$t.err
  echo "== $t" >> gpurun_out/q_summary.txt; grep "psi class" gpurun_out/q_$t.err | tail -3 >> gpurun_out/q_summary.txt
done
