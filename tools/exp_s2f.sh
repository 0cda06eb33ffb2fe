#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py tests/test_gpu_umma.py -x -q > gpurun_out/exp_tests5.log 2>&1
tail -3 gpurun_out/exp_tests5.log
pick() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open(path) if l.startswith('{')][-1])
    k = j.get('kernels_ms_per_step', {})
    print(name, 'ms/step', round(j['ms_per_step'], 4), 'value', round(j['value']), 'e2e', round(j.get('e2e', {}).get('value', 0)), {a: round(b, 4) for a, b in k.items() if b > 0.1})
except Exception as e:
    print(name, 'FAILED', e)
PY
}
B="python bench.py --no-extras --no-cpu-baseline"
for ns in 0 100 400; do
  DRB_WAIT_NS=$ns timeout 200 $B --steps 20 --warmup 5 > gpurun_out/exp5_c3_ns$ns.json 2> gpurun_out/exp5_c3_ns$ns.err; pick c3_wait$ns gpurun_out/exp5_c3_ns$ns.json
done
for ns in 0 100 400; do
  DRB_WAIT_NS=$ns DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp5_c4f_ns$ns.json 2> gpurun_out/exp5_c4f_ns$ns.err; pick c4f_wait$ns gpurun_out/exp5_c4f_ns$ns.json
done
DRB_SCORE_DEBUG=8 DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 3 --warmup 3 > gpurun_out/exp5_c4f_dbg8.json 2>&1; pick c4f_dbg8 gpurun_out/exp5_c4f_dbg8.json
