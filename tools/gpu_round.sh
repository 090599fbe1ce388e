#!/bin/bash
# One GPU-box session: tests, benches, ncu evidence.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python tools/microbench.py --what setops > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; cat gpurun_out/microbench.jsonl; tail -5 gpurun_out/microbench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:setop_pipe_kernel -c 2 -o gpurun_out/setop_prof -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_setop.err; tail -3 gpurun_out/ncu_setop.err
ls -la gpurun_out
