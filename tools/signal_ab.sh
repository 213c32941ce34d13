#!/bin/bash
# usage: signal_ab.sh N — the C2 bench at N GPUs with frame signals through the shared framebuffers, then with the NCCL barrier
N=$1
for mode in signals nccl; do
  if [ $mode = nccl ]; then export COH_NCCL_BARRIER=1; else unset COH_NCCL_BARRIER; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/s_${mode}_n$N.err | grep "^{" > gpurun_out/s_${mode}_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/s_${mode}_n$N.json')); print('$mode N=$N', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], json.dumps(d.get('per_rank'))[:700])"
  tail -3 gpurun_out/s_${mode}_n$N.err
done
