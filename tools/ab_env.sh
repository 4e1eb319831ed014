#!/bin/bash
# bench --quick-ish runs under different environment settings: tools/ab_env.sh tag "VAR=val VAR2=val" "VAR=val" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting python bench.py --no-survey --no-extra --no-cpu --steps 20 --warmup 5 > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_$i.json")); k=j["kernels"]
    print("[$setting] ms_per_step %.3f | " % j["ms_per_step"] + " ".join("%s=%.3f" % (n, k[n]["ms"]) for n in sorted(k) if n.startswith(("agg","proj","heads","graph_build_total"))))
except Exception as e:
    print("[$setting] failed", e)
PY
done
