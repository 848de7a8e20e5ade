#!/bin/bash
# A/B of tunables: each line of VARIANTS is an env assignment string; prints images/s and the top launches.
mkdir -p gpurun_out
IFS=';' read -ra VARS <<< "$VARIANTS"
for v in "${VARS[@]}"; do
  echo "=== variant: $v"
  env $v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-launches gpurun_out/ab.json > gpurun_out/ab_bench.json 2>gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_bench.json")); print("images/s %.1f  ms/step %.3f  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
l=json.load(open("gpurun_out/ab.json")); agg={}
for r in l["launches"]:
    k=(r["kernel"], tuple(r["shape"])[:5]); a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=r["ms"]
print("kernel-sum ms %.3f" % l["total_ms"])
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print("   %-24s %-26s n=%2d ms=%.3f" % (k[0], k[1], a[0], a[1]))
PY
done
