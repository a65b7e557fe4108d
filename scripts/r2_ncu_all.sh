# ncu --set full captures of every hot kernel at HEAD (one GPU; each program runs once without ncu first)
set -x
P="ncu --set full --clock-control none --import-source on"
python scripts/prof.py --what fused --steps 3 > gpurun_out/r2_prof_fused.log 2>&1 && \
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02_ncu_fused.log 2>&1
python scripts/prof.py --what fused_distinct --steps 3 > gpurun_out/r2_prof_fd.log 2>&1 && \
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02_fused_distinct python scripts/prof.py --what fused_distinct --steps 3 > gpurun_out/r02_ncu_fd.log 2>&1
python scripts/prof.py --what pipeline --steps 3 > gpurun_out/r2_prof_pipe.log 2>&1 && \
$P -k regex:'afterstates_kernel|step_kernel' -s 6 -c 2 -o gpurun_out/r02_pipe python scripts/prof.py --what pipeline --steps 3 > gpurun_out/r02_ncu_pipe.log 2>&1
python scripts/prof.py --what rollout_random --steps 2 > gpurun_out/r2_prof_rr.log 2>&1 && \
$P -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/r02_rollout_random python scripts/prof.py --what rollout_random --steps 2 > gpurun_out/r02_ncu_rr.log 2>&1
python scripts/prof.py --what rollout_greedy --steps 2 > gpurun_out/r2_prof_rg.log 2>&1 && \
$P -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/r02_rollout_greedy python scripts/prof.py --what rollout_greedy --steps 2 > gpurun_out/r02_ncu_rg.log 2>&1
python scripts/value_prof.py > gpurun_out/r2_prof_value.log 2>&1 && \
$P -k regex:value_rows_kernel -s 2 -c 1 -o gpurun_out/r02_value python scripts/value_prof.py > gpurun_out/r02_ncu_value.log 2>&1
cat gpurun_out/r2_prof_*.log
