set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; tail -3 gpurun_out/bench_r01f.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r01f.json 2> gpurun_out/bench_ref_r01f.err; cat gpurun_out/bench_ref_r01f.json | cut -c1-300
export WQ_BENCH_GENES=40000 WQ_BENCH_READS=2000000 WQ_BENCH_CPU_SAMPLE=100000
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"wq_k_extend|wq_k_seed|wq_k_resolve|wq_k_em_psi|wq_k_em_tpm|wq_k_beta_ci|wq_k_psi_ci" -s 14 -c 14 -o gpurun_out/prof_r01f python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -4
