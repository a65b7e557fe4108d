set -x
export TPL_PAIR=1
P="ncu --set full --clock-control none --import-source on"
python scripts/prof.py --what fused --steps 3 > gpurun_out/r2b_prof_fused.log 2>&1 && \
$P -k regex:pair_kernel -s 3 -c 1 -o gpurun_out/r02b_pair_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02b_ncu_fused.log 2>&1
$P -k regex:pair_kernel -s 3 -c 1 -o gpurun_out/r02b_pair_as python scripts/prof.py --what pipeline --steps 3 > gpurun_out/r02b_ncu_as.log 2>&1
cat gpurun_out/r2b_prof_*.log
