#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py -x -q > gpurun_out/exp_tests4.log 2>&1
tail -3 gpurun_out/exp_tests4.log
pick() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open(path) if l.startswith('{')][-1])
    k = j.get('kernels_ms_per_step', {})
    print(name, 'ms/step', round(j['ms_per_step'], 4), 'value', round(j['value']), 'e2e', round(j.get('e2e', {}).get('value', 0)), {a: round(b, 3) for a, b in k.items() if b > 0.1})
except Exception as e:
    print(name, 'FAILED', e)
PY
}
B="python bench.py --no-extras --no-cpu-baseline"
for cfg in "16384 3 1024" "16384 4 512" "18944 3 1024" "18944 4 512"; do
  set -- $cfg
  DRB_BENCH_SCORE_BATCH=$1 DRB_TOPK_GROWTH=$2 DRB_TOPK_NS=$3 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp4_c4f_$1_$2_$3.json 2> gpurun_out/exp4_c4f_$1_$2_$3.err
  pick c4f_sb$1_g$2_ns$3 gpurun_out/exp4_c4f_$1_$2_$3.json
done
DRB_SCORE_DEBUG=1 timeout 300 $B --workload c4_full --steps 3 --warmup 3 > gpurun_out/exp4_c4f_dbg1.json 2>&1; pick dbg1 gpurun_out/exp4_c4f_dbg1.json
