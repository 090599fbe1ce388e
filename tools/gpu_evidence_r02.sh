#!/bin/bash
# The evidence session of round 2 (ONE GPU): whole GPU suite, bench (both arms), microbenchmarks incl. the cub yardstick,
# C4, ncu launch list and full captures.  Everything lands in gpurun_out/; tools/make_profiles_r02.py turns it into profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu_info.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python tools/microbench.py --what sort1,cub,pairs,kmers,count,fold,minimizer > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; cat gpurun_out/microbench.jsonl; tail -3 gpurun_out/microbench.err
timeout 600 python tools/run_c4.py > gpurun_out/c4.json 2> gpurun_out/c4.err; tail -2 gpurun_out/c4.json; tail -2 gpurun_out/c4.err
timeout 300 python tools/exp_ops.py --ops inter,diff,union --reps 5 --variants "default:;chain_walk:UKM_NFILTER=0;rows_union:UKM_NUNION=1" > gpurun_out/exp_ops.jsonl 2> gpurun_out/exp_ops.err; cut -c1-300 gpurun_out/exp_ops.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/ncu_launches.err; tail -2 gpurun_out/ncu_launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_kernel -s 1 -c 1 -o gpurun_out/nway_union3_prof -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2> gpurun_out/ncu_nway.err; tail -2 gpurun_out/ncu_nway.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nfilter_kernel -s 1 -c 1 -o gpurun_out/nfilter_prof -f python tools/exp_ops.py --ops inter --reps 1 > /dev/null 2> gpurun_out/ncu_nfilter.err; tail -2 gpurun_out/ncu_nfilter.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nway_kernel -s 1 -c 1 -o gpurun_out/nway_union_prof -f python tools/exp_ops.py --ops union --reps 1 > /dev/null 2> gpurun_out/ncu_nwayu.err; tail -2 gpurun_out/ncu_nwayu.err
ls -la gpurun_out | head -60
