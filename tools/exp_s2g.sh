#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/exp_tests6.log 2>&1
tail -3 gpurun_out/exp_tests6.log
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
for kb in 16 32; do
  DRB_SCORE_KB=$kb DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp6_c4f_kb$kb.json 2> gpurun_out/exp6_c4f_kb$kb.err; pick c4f_kb$kb gpurun_out/exp6_c4f_kb$kb.json
done
DRB_TOPK_GROWTH=4 DRB_TOPK_NS=512 DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp6_c4f_g4.json 2>&1; pick c4f_g4_ns512 gpurun_out/exp6_c4f_g4.json
for d in 1 8; do
DRB_SCORE_DEBUG=$d DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 3 --warmup 3 > gpurun_out/exp6_c4f_dbg$d.json 2>&1; pick c4f_dbg$d gpurun_out/exp6_c4f_dbg$d.json
done
