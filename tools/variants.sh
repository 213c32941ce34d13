for v in "" w2_c3 w2_c4 w4_c3 w4_c4 w1_c4; do
  if [ -z "$v" ]; then unset COH_LIB_PATH; else export COH_LIB_PATH=$PWD/coherence_renderer_b200/libcoh_$v.so; fi
  echo "== ${v:-default}"; python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['binning_ms'])"
done
