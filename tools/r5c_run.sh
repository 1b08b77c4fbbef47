mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_tfm_train_gpu.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -2
for occ in 0 1; do for wl in cfg3 cfg5; do
DOF_ROWS_OCC2=$occ timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-secondary | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ2=$occ $wl', round(d['value']), round(d['ms_per_step'],2), {k: round(v['ms_per_step'],2) for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:5]})"
done; done
