for ab in 0 4; do
COH_AB=$ab python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$ab bench.py --gpus 4 --steps 200 --warmup 20 --no-cpu-baseline 2> gpurun_out/d4.err | grep "^{" > gpurun_out/d4.json; python -c "
import json; d=json.load(open('gpurun_out/d4.json')); print('N=4 display ab=$ab', d['ms_per_step'], d['e2e']['ms_per_step'], [(round(r['raster_ms'],4), round(r['barrier_wait_ms'],4)) for r in d['per_rank']])"; tail -2 gpurun_out/d4.err
done
python -m pytest tests -m gpu -x -q -k "peer or band or multi or abi" 2>&1 | tail -3
