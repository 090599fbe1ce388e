#!/bin/bash
# A/B of the peer-pull exchange at N=4 (where it bounds the step).  Variant = "NATIVE STREAMS SPLIT SRC_STREAMS"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4"
i=0
while read -r nat st sp ss; do
  [ -z "$nat" ] && continue
  i=$((i+1)); tag="n4_nat${nat}_s${st}_p${sp}_q${ss}"
  UKM_PULL_NATIVE=$nat UKM_PULL_STREAMS=$st UKM_PULL_SPLIT=$sp UKM_PULL_SRC_STREAMS=$ss timeout 150 $TR --master-port $((29520+i)) bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  grep -v "CudaIPCTypes\|^\*\*\*\|NCCL version" gpurun_out/$tag.err | tail -2
  python -c "
import json,sys
d=json.loads(open('gpurun_out/$tag.json').read().strip().splitlines()[-1])
x=d['exchange']
print('$tag', round(d['ms_per_step'],3), round(d['roofline']['share_of_step'],3), d['check']['union_full_digest_exact'], {k:(round(x[k]['span_ms'],2), round(x[k]['GBps_over_span'],1)) for k in ('under_kernels','gpu_idle')})
"
done <<< "${VARIANTS:-1 6 1 1
1 12 2 1
0 6 1 2}"
