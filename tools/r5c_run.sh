mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r5y_tests.log 2>&1
tail -4 gpurun_out/r5y_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5y_smoke.log 2>&1; tail -5 gpurun_out/r5y_smoke.log
timeout 900 python bench.py > gpurun_out/r5y_bench.json 2> gpurun_out/r5y_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r5y_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:10]:
    print("  %-16s n=%4.0f %8.3f ms %5.1f%% tf=%s gbs=%s" % (k, v["launches_per_step"], v["ms_per_step"], 100 * v["share"], round(v.get("tflops", 0), 1), round(v.get("gbs", 0))))
for s in d["secondary"]:
    print(s.get("workload","")[:30], round(s.get("value",0)), s.get("ms_per_step"), s.get("error"))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r5y_bench_reference_arm.json 2>&1; tail -c 600 gpurun_out/r5y_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r5y_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r5y_launches.csv > gpurun_out/r5y_ncu_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_(fwd|bwdw)_tc" -s 33 -c 11 -f -o /tmp/r5y_gru python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r5y_ncu_gru.log 2>&1
python tools/ncu_summary.py /tmp/r5y_gru.ncu-rep > gpurun_out/r5y_ncu_full_gru_kernels.txt
rm -f gpurun_out/r5y_launches.csv
ls -la gpurun_out | tail -12
