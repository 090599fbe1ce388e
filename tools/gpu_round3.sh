#!/bin/bash
# GPU session: N-way union after the partition / search changes -- parity, A/B of shapes, launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nway.py -q -x --timeout=180 -p no:cacheprovider > gpurun_out/pytest_nway.log 2>&1; NW=$?; tail -15 gpurun_out/pytest_nway.log
timeout 600 python tools/exp_nway.py > gpurun_out/exp_nway.jsonl 2> gpurun_out/exp_nway.err; cat gpurun_out/exp_nway.jsonl; tail -5 gpurun_out/exp_nway.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_nway.csv python tools/exp_nway.py --cfgs 0 > /dev/null 2> gpurun_out/ncu_launches.err; tail -3 gpurun_out/ncu_launches.err
ls -la gpurun_out
