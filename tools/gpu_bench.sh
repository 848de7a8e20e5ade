#!/bin/bash
# bench + launch list + (optional) ncu full capture of a few kernels; keeps gpurun_out/ small (CSV summaries only).
mkdir -p gpurun_out
R=${ROUND:-r01}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_$R.csv &
SMI=$!
timeout 300 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS} --dump-launches gpurun_out/launches_events_$R.json > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc $?"; cat gpurun_out/bench_$R.json; tail -5 gpurun_out/bench_$R.err
kill $SMI
python - <<PY
import json
d=json.load(open("gpurun_out/launches_events_$R.json"))
print("total_ms", d["total_ms"])
agg={}
for r in d["launches"]:
    k=(r["kernel"], tuple(r["shape"]))
    a=agg.setdefault(k,[0,0.0,0.0]); a[0]+=1; a[1]+=r["ms"]; a[2]+=r["gflop"]
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:40]:
    print("%-28s %-28s n=%2d ms=%7.3f gflop=%8.1f TF/s=%6.1f" % (k[0], k[1], a[0], a[1], a[2], a[2]/a[1] if a[1] else 0))
PY
if [ -n "$NCU_LIST" ]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"igemm|maxpool|upsample|edge_|nms_|topk_|pack_image|sigmoid|head_fused|rows_conv" -c 700 --csv --log-file gpurun_out/launches_ncu_$R.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$R.log 2>&1
echo "ncu list rc $?"
fi
if [ -n "$NCU_FULL" ]; then
timeout 600 ncu --set full --kernel-name regex:"igemm|nms_|topk_|head_fused|rows_conv" --clock-control none --import-source on --profile-from-start off -f -o /tmp/prof_$R python tools/profile_kernels.py > gpurun_out/ncu_full_$R.log 2>&1
echo "ncu full rc $?"; tail -3 gpurun_out/ncu_full_$R.log
ncu -i /tmp/prof_$R.ncu-rep --page raw --csv > gpurun_out/prof_raw_$R.csv 2>/dev/null
ncu -i /tmp/prof_$R.ncu-rep --page details --csv > gpurun_out/prof_details_$R.csv 2>/dev/null
sz=$(stat -c %s /tmp/prof_$R.ncu-rep); echo "rep size $sz"
if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_$R.ncu-rep gpurun_out/; fi
fi
ls -la gpurun_out/; du -sh gpurun_out
