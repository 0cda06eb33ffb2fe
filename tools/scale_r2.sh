#!/bin/bash
# Round-2 multi-GPU runs on ONE box (gpurun --gpus 8): configs[4] item-sharded, the c3 weak-scaling curve, strong scaling,
# item-sharded c3, DMF and full-catalog ranking at 8 GPUs.  Each line lands in gpurun_out/scale_r2_*.json.
cd "${GRAFT_REPO_ROOT:-.}"
NG=${NG:-8}
run() {  # name, nproc, args...
  local name=$1 n=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n --no-cpu-baseline --no-extras "$@" 2> gpurun_out/scale_r2_$name.err | grep '^{' > gpurun_out/scale_r2_$name.json
  python - <<PY
import json
try:
    j = json.load(open('gpurun_out/scale_r2_$name.json'))
    print('$name', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'], 4), 'e2e', round(j.get('e2e', {}).get('value', 0)),
          'parity', j.get('dp_parity', {}).get('data', {}).get('ok'), j.get('dp_parity', {}).get('items', {}).get('ok'),
          'mem', j['config'].get('memory_per_rank_gb'))
except Exception as e:
    print('$name', 'FAILED', e)
    print(open('gpurun_out/scale_r2_$name.err').read()[-1500:])
PY
}
run c5_n$NG $NG --workload c5 --steps 10 --warmup 3
run c3_n$NG $NG
if [ "$NG" -ge 4 ]; then run c3_n4 4 --no-dp-parity; fi
if [ "$NG" -ge 2 ] && [ "$NG" -ne 2 ]; then run c3_n2 2 --no-dp-parity; fi
run c3_strong_n$NG $NG --scaling strong --no-dp-parity
run c3_items_n$NG $NG --parallel items --no-dp-parity
run c2_n$NG $NG --workload c2
run c4full_n$NG $NG --workload c4_full
