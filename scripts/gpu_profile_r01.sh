set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
export WQ_BENCH_GENES=40000 WQ_BENCH_READS=2000000 WQ_BENCH_CPU_SAMPLE=100000
python bench.py --steps 2 --warmup 1 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wq_k_extend -s 1 -c 2 -o gpurun_out/prof_extend_r01 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu2.log 2>&1
python bench.py --steps 2 --warmup 1 > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wq_k_seed -s 1 -c 1 -o gpurun_out/prof_seed_r01 python bench.py --steps 2 --warmup 1 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
