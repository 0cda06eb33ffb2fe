set -x
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
for d in 0 2; do
DRB_LOSS_DEBUG=$d timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/dbg_$d.json 2> gpurun_out/dbg_$d.err
python - <<PY
import json
l=[x for x in open('gpurun_out/dbg_$d.json') if x.startswith('{')]
j=json.loads(l[-1]); print('DBG',$d, j['ms_per_step'], j['value'], j['e2e']['value'], {k:v for k,v in j.get('kernels_ms_per_step',{}).items()})
PY
done
