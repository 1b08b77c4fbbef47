mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "weight_gradients" > gpurun_out/r5g_tests_bwdw.log 2>&1
tail -3 gpurun_out/r5g_tests_bwdw.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r5g_bench.json 2> gpurun_out/r5g_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r5g_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:8]:
    print("  %-16s n=%4.0f %8.3f ms %5.1f%% tf=%s gbs=%s" % (k, v["launches_per_step"], v["ms_per_step"], 100 * v["share"], round(v.get("tflops", 0), 1), round(v.get("gbs", 0))))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_bwdw" -s 16 -c 2 -f -o /tmp/r5g python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r5g_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r5g.ncu-rep > gpurun_out/r5g_ncu_full_gru_bwdw.txt
ncu -i /tmp/r5g.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r5g_source_sass.csv.gz
