#!/bin/bash
# One GPU-box session for the N-way union: its parity tests first (the rest of the session falls back to the two-way
# tree if they fail), the whole GPU suite, the A/B of the union paths, the bench, ncu evidence.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu_info.csv 2>&1; cat gpurun_out/gpu_info.csv
timeout 900 python -m pytest tests/test_gpu_nway.py -q -x --timeout=180 -p no:cacheprovider > gpurun_out/pytest_nway.log 2>&1; NW=$?; tail -25 gpurun_out/pytest_nway.log
if [ $NW -ne 0 ]; then export UKM_NWAY=0; echo "NWAY DISABLED FOR THE REST OF THE SESSION"; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --deselect tests/test_gpu_nway.py > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
if [ $NW -eq 0 ]; then
  timeout 600 python tools/exp_nway.py --h2d > gpurun_out/exp_nway.jsonl 2> gpurun_out/exp_nway.err; cat gpurun_out/exp_nway.jsonl; tail -5 gpurun_out/exp_nway.err
fi
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 4000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -3 gpurun_out/ncu_launches.err
if [ $NW -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_union_kernel -s 1 -c 1 -o gpurun_out/nway_union_prof -f python tools/exp_nway.py --cfgs 0 > /dev/null 2> gpurun_out/ncu_nway.err; tail -3 gpurun_out/ncu_nway.err
fi
ls -la gpurun_out
