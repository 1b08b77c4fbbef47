mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_tfm_train_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r5o_tests.log 2>&1
tail -4 gpurun_out/r5o_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5o_bench.json 2> gpurun_out/r5o_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r5o_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for s in d["secondary"]:
    print(s.get("workload","")[:30], round(s.get("value",0)), s.get("ms_per_step"), s.get("error"))
    km=s["kernels_ms"]; print("    ", sorted(km.items(), key=lambda kv:-kv[1])[:6])
PY
