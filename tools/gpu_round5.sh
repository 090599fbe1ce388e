#!/bin/bash
# GPU session: union with per-level de-duplication -- parity, A/B, bench, launch list, ncu captures.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nway.py -q -x --timeout=240 -p no:cacheprovider > gpurun_out/pytest_nway.log 2>&1; tail -5 gpurun_out/pytest_nway.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider --deselect tests/test_gpu_nway.py > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/exp_nway.py --cfgs off,0,1,2,4 > gpurun_out/exp_nway.jsonl 2> gpurun_out/exp_nway.err; cat gpurun_out/exp_nway.jsonl; tail -5 gpurun_out/exp_nway.err
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 3000 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -3 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_kernel -s 1 -c 1 -o gpurun_out/nway_union_prof -f python tools/exp_nway.py --cfgs 0 --only union > /dev/null 2> gpurun_out/ncu_nway.err; tail -3 gpurun_out/ncu_nway.err
UKM_NWAY_FILTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_kernel -s 1 -c 1 -o gpurun_out/nway_filter_prof -f python tools/exp_nway.py --cfgs 0 --only inter > /dev/null 2> gpurun_out/ncu_nwayf.err; tail -3 gpurun_out/ncu_nwayf.err
ls -la gpurun_out
