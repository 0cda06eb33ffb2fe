#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/exp_tests7.log 2>&1
tail -3 gpurun_out/exp_tests7.log
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
for w in 1 0; do
  DRB_SELECT_WARP=$w DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp7_c4f_w$w.json 2> gpurun_out/exp7_c4f_w$w.err; pick c4f_warpsel$w gpurun_out/exp7_c4f_w$w.json
done
DRB_TOPK_NS=2048 DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp7_c4f_ns2048.json 2>&1; pick c4f_ns2048 gpurun_out/exp7_c4f_ns2048.json
DRB_TOPK_NS=1024 DRB_TOPK_GROWTH=2 DRB_BENCH_SCORE_BATCH=18944 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp7_c4f_g2.json 2>&1; pick c4f_g2 gpurun_out/exp7_c4f_g2.json
