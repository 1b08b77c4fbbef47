mkdir -p gpurun_out
python tools/prof_gru.py 32 32 57344 all 1 > gpurun_out/r5u_fwd_timeline_h32.txt 2>&1; python tools/prof_gru.py 16 64 57344 all 1 > gpurun_out/r5u_fwd_timeline_h16.txt 2>&1
grep "fused gru\|cycles per step" gpurun_out/r5u_fwd_timeline_h32.txt gpurun_out/r5u_fwd_timeline_h16.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r5u_tests.log 2>&1
tail -4 gpurun_out/r5u_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r5u_bench.json 2> gpurun_out/r5u_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r5u_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:10]:
    print("  %-16s n=%4.0f %8.3f ms %5.1f%% tf=%s gbs=%s" % (k, v["launches_per_step"], v["ms_per_step"], 100 * v["share"], round(v.get("tflops", 0), 1), round(v.get("gbs", 0))))
print("bytes/step GB", sum(v["bytes_per_step"] for v in d["kernels"].values())/1e9, "roofline", {k: d["roofline"][k] for k in ("kernel","bound","achieved","peak","frac")})
for s in d["secondary"]:
    print(s.get("workload","")[:30], round(s.get("value",0)), s.get("ms_per_step"), s.get("error"))
PY
