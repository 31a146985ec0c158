set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_r01d.log; cat gpurun_out/pytest_r01d.log
WQ_PROFILE=1 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -20 gpurun_out/bench_r01d.err; cat gpurun_out/bench_r01d.json
export WQ_BENCH_GENES=40000 WQ_BENCH_READS=2000000 WQ_BENCH_CPU_SAMPLE=100000
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"wq_k_extend|wq_k_seed|wq_k_resolve|wq_k_em_psi|wq_k_em_tpm" -s 6 -c 6 -o gpurun_out/prof_r01d python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -4
