set -x
export TPL_PAIR=0
P="ncu --set full --clock-control none --import-source on"
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02b_old_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02b_ncu_oldfused.log 2>&1
tail -2 gpurun_out/r02b_ncu_oldfused.log
