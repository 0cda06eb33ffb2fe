#!/bin/bash
# Session experiments (one GPU): correctness of the staged top-k / new loss epilogue, then A/B timings.
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_baseline_shapes.py tests/test_gpu_umma.py tests/test_gpu_fullsize.py -x -q > gpurun_out/exp_tests.log 2>&1
tail -5 gpurun_out/exp_tests.log
pick() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open(path) if l.startswith('{')][-1])
    k = j.get('kernels_ms_per_step', {})
    print(name, 'ms/step', round(j['ms_per_step'], 4), 'value', round(j['value']), 'e2e', round(j.get('e2e', {}).get('value', 0)),
          {a: round(b, 4) for a, b in k.items() if b > 0.05 or 'loss' in a})
except Exception as e:
    print(name, 'FAILED', e)
PY
}
B="python bench.py --no-extras --no-cpu-baseline"
for dbg in 0 64 2 4 6 16; do
  DRB_LOSS_DEBUG=$dbg timeout 200 $B --steps 20 --warmup 5 > gpurun_out/exp_c3_dbg$dbg.json 2> gpurun_out/exp_c3_dbg$dbg.err
  pick c3_dbg$dbg gpurun_out/exp_c3_dbg$dbg.json
done
for cfg in "16384 3 1024" "4096 3 1024" "8192 3 1024" "16384 2 1024" "16384 4 1024" "16384 3 512" "16384 3 2048" "16384 1000 2048"; do
  set -- $cfg
  DRB_BENCH_SCORE_BATCH=$1 DRB_TOPK_GROWTH=$2 DRB_TOPK_NS=$3 timeout 300 $B --workload c4_full --steps 5 --warmup 3 > gpurun_out/exp_c4f_$1_$2_$3.json 2> gpurun_out/exp_c4f_$1_$2_$3.err
  pick c4f_sb$1_g$2_ns$3 gpurun_out/exp_c4f_$1_$2_$3.json
done
