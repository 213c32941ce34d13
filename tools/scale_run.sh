#!/bin/bash
# usage: scale.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/r2_bench_n$N.err | grep "^{" > gpurun_out/r2_bench_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print('N=$N', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], json.dumps(d.get('per_rank'))[:600])"
tail -3 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/c3_bands.py 2> gpurun_out/r2_c3_n$N.err | grep "^{" > gpurun_out/r2_c3_n$N.json; cat gpurun_out/r2_c3_n$N.json; tail -3 gpurun_out/r2_c3_n$N.err
