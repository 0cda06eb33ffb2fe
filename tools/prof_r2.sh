#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun; one GPU).  Outputs land in gpurun_out/, summarised into
# profiles/r2_* by profiles/summarize.py.  A number printed by a run under ncu is never a bench value.
set -x
cd "${GRAFT_REPO_ROOT:-.}"
B="python bench.py --no-extras --no-cpu-baseline --no-flush"
# 1. every launch of two c3 steps with its device time (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2_c3.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
# 2. full metrics of the top c3 kernels (one launch each, after 4 warm steps)
ncu --set full --clock-control none --import-source on -k regex:'k_umma_cdae_loss|k_umma_gemm|k_adam|k_scatter_chunks|k_gather_chunks' -s 24 -c 6 -f -o gpurun_out/prof_r2_c3 $B --steps 2 --warmup 3 > /dev/null 2>&1
# 3. full-catalog top-k: launch list + full metrics of the filter / select kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_topk.csv python tools/prof_topk.py 18944 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_umma_score_filter|k_select_lists|k_select_tau_warp' -s 8 -c 8 -f -o gpurun_out/prof_r2_topk python tools/prof_topk.py 18944 3 > /dev/null 2>&1
# 4. DMF step (config 2), kernels launched directly so that every node is visible
DRB_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_dmf.csv $B --workload c2 --steps 50 --warmup 3 > /dev/null 2>&1
DRB_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'k_gather|k_scatter|k_dmf|k_adam' -s 50 -c 5 -f -o gpurun_out/prof_r2_dmf $B --workload c2 --steps 50 --warmup 3 > /dev/null 2>&1
# 5. sampled-candidate ranking (config 4): gather + rank kernels
ncu --set full --clock-control none --import-source on -k regex:'k_rank_candidates|k_gather' -s 40 -c 2 -f -o gpurun_out/prof_r2_rank $B --workload c4_sampled --steps 3 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_*.csv
