#!/bin/bash
# C3 bands at N GPUs: gather to the display rank / to every rank, walker work items of 16 / 4 rows
N=${1:-4}
for g in display all; do for wh in 0 4; do
echo "gather=$g walk_h=$wh"
COH_GATHER=$g COH_WALK_H=$wh python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/c3_bands.py 2>&1 | grep "^{\|Error\|error" | head -5
done; done
