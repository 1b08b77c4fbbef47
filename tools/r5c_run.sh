mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8 > gpurun_out/r5n_topo.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r5n_tests_mgpu.log 2>&1
tail -12 gpurun_out/r5n_tests_mgpu.log
for ar in peer nccl; do
DOF_ALLREDUCE=$ar timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > gpurun_out/r5n_bench_n2_$ar.json 2> gpurun_out/r5n_bench_n2_$ar.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r5n_bench_n2_$ar.json").read().strip().splitlines()[-1])
    print("$ar", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("allreduce"))
    print({k: round(v["ms_per_step"],4) for k,v in d["kernels"].items() if "peer" in k or "adam" in k})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r5n_bench_n2_$ar.err").read()[-3000:])
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', round(d['value']), round(d['ms_per_step'],3))"
