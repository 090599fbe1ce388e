#!/bin/bash
# One GPU-box session: tests, benches, ncu evidence.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --universe 1e8 --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 900 python tools/microbench.py > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; cat gpurun_out/microbench.jsonl; tail -5 gpurun_out/microbench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -3 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:setop_kernel -c 3 -o gpurun_out/setop_prof -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_full.err; tail -3 gpurun_out/ncu_full.err
ls -la gpurun_out
