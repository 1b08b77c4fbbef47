mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 --no-secondary > gpurun_out/r6f_bench_n$n.json 2> gpurun_out/r6f_bench_n$n.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r6f_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("allreduce", {}).get("kind"))
except Exception as e:
    print("N=$n parse failed", e); print(open("gpurun_out/r6f_bench_n$n.err").read()[-2500:])
PY
done
