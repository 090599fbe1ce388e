#!/bin/bash
# A/B of the peer-pull exchange at N=4 (where it bounds the step): bench.py --gpus 4 under variants of
# "copy streams of the puller / pieces per pulled slice / streams per owner device" (unikmer_b200/dist.py).
# Results of round 2: profiles/r02_exp_pull_n4.md.   gpurun --gpus 4 -- 'bash tools/exp_pull_n4.sh'
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
i=0
while read -r st sp ss; do
  [ -z "$st" ] && continue
  i=$((i+1)); tag="n4_s${st}_p${sp}_q${ss}"
  UKM_PULL_STREAMS=$st UKM_PULL_SPLIT=$sp UKM_PULL_SRC_STREAMS=$ss timeout 150 $TR --master-port $((29520+i)) bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  grep -v "CudaIPCTypes\|^\*\*\*\|NCCL version\|OMP_NUM" gpurun_out/$tag.err | tail -2
  python -c "
import json,sys
d=json.loads(open('gpurun_out/$tag.json').read().strip().splitlines()[-1])
x=d['exchange']
print('$tag', round(d['ms_per_step'],3), round(d['roofline']['share_of_step'],3), d['check']['union_full_digest_exact'], {k:(round(x[k]['span_ms'],2), round(x[k]['GBps_over_span'],1)) for k in ('under_kernels','gpu_idle')})
"
done <<< "${VARIANTS:-8 1 1
8 1 2
12 2 2}"
