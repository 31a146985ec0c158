#!/bin/bash
# round 2, GPU call B: parity tests after the event-stage / PSI changes, config-2 bench with timeline, config 1 and 3 quick runs
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/b_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
WQ_PROFILE=2 timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_bench.err
timeout 600 python bench.py --config 1 --steps 3 --warmup 2 > gpurun_out/b_bench_c1.json 2> gpurun_out/b_bench_c1.err; echo "rc=$?" >> gpurun_out/b_bench_c1.err
timeout 900 python bench.py --config 3 --reads 4000000 --steps 3 --warmup 2 --cpu-sample 600000 > gpurun_out/b_bench_c3.json 2> gpurun_out/b_bench_c3.err; echo "rc=$?" >> gpurun_out/b_bench_c3.err
timeout 900 python bench.py --index synth40k --steps 3 --warmup 2 --no-parity > gpurun_out/b_bench_synth.json 2> gpurun_out/b_bench_synth.err; echo "rc=$?" >> gpurun_out/b_bench_synth.err
