set -x
for p in 0 1; do
  echo "=== TPL_PAIR=$p"
  TPL_PAIR=$p python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
  TPL_PAIR=$p python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
done
export TPL_PAIR=0
P="ncu --set full --clock-control none --import-source on"
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02d_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02d_ncu_fused.log 2>&1
