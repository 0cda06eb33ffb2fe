#!/bin/bash
# Closing multi-GPU lines on ONE box (gpurun --gpus 8): c3 weak scaling and full-catalog ranking at 8 GPUs.
cd "${GRAFT_REPO_ROOT:-.}"
NG=${NG:-8}
run() {  # name, args...
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $NG --no-cpu-baseline --no-extras "$@" 2> gpurun_out/r2f_$name.err | grep '^{' > gpurun_out/r2f_$name.json
  python - <<PY
import json
try:
    j = json.load(open('gpurun_out/r2f_$name.json'))
    print('$name', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'], 4), 'e2e', round(j.get('e2e', {}).get('value', 0)),
          'parity', j.get('dp_parity', {}).get('data', {}).get('ok'), j.get('dp_parity', {}).get('items', {}).get('ok'))
    print('   ', {a: round(b, 4) for a, b in j.get('kernels_ms_per_step', {}).items()})
except Exception as e:
    print('$name', 'FAILED', e)
    print(open('gpurun_out/r2f_$name.err').read()[-1200:])
PY
}
run bench_c3_n$NG
run bench_c4full_n$NG --workload c4_full --steps 3 --warmup 3
