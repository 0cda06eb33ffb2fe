#!/bin/bash
# Round-2 closing run on one GPU: smoke(), the whole GPU test-suite, the default bench line (all workloads with CPU
# baselines), the reference arm, and the ncu evidence for the final kernels (launch lists = shares of the step; full
# captures of the top kernels).  A number printed by a run under ncu is never a bench value.
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2f_gputests.log 2>&1
tail -4 gpurun_out/r2f_gputests.log
timeout 600 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 300 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
tail -c 400 gpurun_out/r2f_bench_reference.json
B="python bench.py --no-extras --no-cpu-baseline --no-flush"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches_c3.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_umma_cdae_loss|k_umma_gemm|k_adam|k_scatter_chunks|k_gather_chunks' -s 24 -c 6 -f -o gpurun_out/r2f_prof_c3 $B --steps 2 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f_launches_topk.csv python tools/prof_topk.py 18944 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_umma_score_filter|k_select_lists|k_select_tau_warp' -s 8 -c 8 -f -o gpurun_out/r2f_prof_topk python tools/prof_topk.py 18944 3 > /dev/null 2>&1
DRB_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_dmf.csv $B --workload c2 --steps 50 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/r2f_*
