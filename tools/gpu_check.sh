#!/bin/bash
# Standard GPU-box check: tests, bench, launch list.  usage: tools/gpu_check.sh <tag> [ncu-kernel-regex]
TAG=${1:-run}
KRE=${2:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d["clocks"])
    for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"])[:14]:
        print("  %-16s n=%4.0f %8.3f ms %5.1f%% tf=%s gbs=%s" % (k, v["launches_per_step"], v["ms_per_step"], 100 * v["share"], round(v.get("tflops", 0), 1), round(v.get("gbs", 0))))
except Exception as e:
    print("bench parse failed", e)
PY
if [ -n "$KRE" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s ${3:-6} -c ${4:-2} -f -o gpurun_out/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
  ls -la gpurun_out/${TAG}_prof* 2>/dev/null
fi
