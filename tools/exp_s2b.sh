#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
pick() { python - "$1" "$2" <<'PY'
import json, sys
name, path = sys.argv[1], sys.argv[2]
try:
    j = json.loads([l for l in open(path) if l.startswith('{')][-1])
    k = j.get('kernels_ms_per_step', {})
    print(name, 'ms/step', round(j['ms_per_step'], 4), 'value', round(j['value']), {a: round(b, 3) for a, b in k.items() if b > 0.3})
except Exception as e:
    print(name, 'FAILED', e)
PY
}
B="python bench.py --no-extras --no-cpu-baseline"
for dbg in 1 8 2 4 12; do
  DRB_SCORE_DEBUG=$dbg timeout 300 $B --workload c4_full --steps 3 --warmup 3 > gpurun_out/exp_c4f_dbg$dbg.json 2> gpurun_out/exp_c4f_dbg$dbg.err
  pick c4f_dbg$dbg gpurun_out/exp_c4f_dbg$dbg.json
done
