set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -3 gpurun_out/bench_r01e.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01e.json 2> gpurun_out/bench_ref_r01e.err; cat gpurun_out/bench_ref_r01e.json | cut -c1-400
export WQ_BENCH_GENES=40000 WQ_BENCH_READS=2000000 WQ_BENCH_CPU_SAMPLE=100000
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"wq_k_extend|wq_k_seed|wq_k_resolve|wq_k_em_psi|wq_k_em_tpm|wq_k_merge" -s 9 -c 9 -o gpurun_out/prof_r01e python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -4
