#!/bin/bash
# The evidence session of a round: whole GPU suite, bench (both arms), microbenchmarks, ncu launch list and full captures.
# Everything lands in gpurun_out/; tools/make_profiles.py turns it into profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu_info.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 4500 gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python tools/microbench.py --what sort1,pairs,kmers,count,fold,minimizer > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; cat gpurun_out/microbench.jsonl; tail -5 gpurun_out/microbench.err
timeout 600 python tools/exp_nway.py --cfgs off,0 --h2d > gpurun_out/exp_nway.jsonl 2> gpurun_out/exp_nway.err; cat gpurun_out/exp_nway.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -3 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_kernel -s 1 -c 1 -o gpurun_out/nway_union_prof -f python tools/exp_nway.py --cfgs 0 --only union > /dev/null 2> gpurun_out/ncu_nway.err; tail -3 gpurun_out/ncu_nway.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:setop_pipe_kernel -s 2 -c 2 -o gpurun_out/setop_pipe_prof -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_setop.err; tail -3 gpurun_out/ncu_setop.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:setop_search_kernel -s 4 -c 1 -o gpurun_out/setop_search_prof -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_search.err; tail -3 gpurun_out/ncu_search.err
ls -la gpurun_out
