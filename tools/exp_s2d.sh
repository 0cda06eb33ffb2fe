#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 700 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py tests/test_gpu_umma.py tests/test_gpu_umma_gemm.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q > gpurun_out/exp_tests3.log 2>&1
tail -5 gpurun_out/exp_tests3.log
pick() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open(path) if l.startswith('{')][-1])
    k = j.get('kernels_ms_per_step', {})
    print(name, 'ms/step', round(j['ms_per_step'], 4), 'value', round(j['value']), 'e2e', round(j.get('e2e', {}).get('value', 0)), {a: round(b, 4) for a, b in k.items() if b > 0.02})
except Exception as e:
    print(name, 'FAILED', e)
PY
}
B="python bench.py --no-extras --no-cpu-baseline"
timeout 200 $B --steps 20 --warmup 5 > gpurun_out/exp3_c3.json 2> gpurun_out/exp3_c3.err; pick c3 gpurun_out/exp3_c3.json
for dbg in 4 6; do DRB_LOSS_DEBUG=$dbg timeout 200 $B --steps 20 --warmup 5 > gpurun_out/exp3_c3_dbg$dbg.json 2>&1; pick c3_dbg$dbg gpurun_out/exp3_c3_dbg$dbg.json; done
timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp3_c4f.json 2> gpurun_out/exp3_c4f.err; pick c4f gpurun_out/exp3_c4f.json
DRB_SCORE_DEBUG=8 timeout 300 $B --workload c4_full --steps 3 --warmup 3 > gpurun_out/exp3_c4f_dbg8.json 2>&1; pick c4f_dbg8 gpurun_out/exp3_c4f_dbg8.json
