#!/bin/bash
# usage (on the GPU box): tools/ab_run.sh variant1 variant2 ...   -> one line per variant: step ms + top ops
mkdir -p gpurun_out
for v in "$@"; do
  lib=dagr_b200/build/variants/$v.so
  [ "$v" = "main" ] && lib=dagr_b200/libdagr_b200.so
  DAGR_B200_LIB=$lib timeout 200 python bench.py --headline-only --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/ab_{v}.json"))
    po = j["roofline"]["per_op_ms"]
    print(v, f"step {j['ms_per_step']:.3f} ms  e2e {j['e2e']['ms_per_step']:.3f}", " ".join(f"{k}={x:.3f}" for k, x in list(po.items())[:3]))
except Exception as e:
    print(v, "FAILED", e)
PY
done
