#!/bin/bash
# One GPU-box session: tests, benches, ncu evidence.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-e2e --no-cpu > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; tail -c 2500 gpurun_out/bench_dev.json; tail -5 gpurun_out/bench_dev.err
timeout 1200 python tools/microbench.py --what setops > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; cat gpurun_out/microbench.jsonl; tail -5 gpurun_out/microbench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:setop_pipe_kernel -c 2 -o gpurun_out/setop_prof -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_setop.err; tail -3 gpurun_out/ncu_setop.err
timeout 300 python tools/exp_merge.py > gpurun_out/exp_merge.jsonl 2>&1; cat gpurun_out/exp_merge.jsonl
ls -la gpurun_out
