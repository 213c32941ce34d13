#!/bin/bash
# Bench the default library and every experiment build (make -C coherence_renderer_b200/csrc variant NAME=x EXTRA=...)
# named on the command line:  tools/variants.sh a b c
for v in "" "$@"; do
  if [ -z "$v" ]; then unset COH_LIB_PATH; else export COH_LIB_PATH=$PWD/coherence_renderer_b200/libcoh_$v.so; fi
  echo "== ${v:-default}"
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/frame', round(d['ms_per_step'],4), 'walker', round(d['roofline']['kernel_ms'],4), 'binning', round(d['roofline']['binning_ms'],4))"
done
