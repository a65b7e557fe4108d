set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_tile or fused_step_observe or rollout_greedy or library_loaded or afterstates_vs_oracle" > gpurun_out/r2b4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b4_pytest.log
for p in 0 1; do
  echo "=== TPL_PAIR=$p"
  TPL_PAIR=$p python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
  TPL_PAIR=$p python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
  TPL_PAIR=$p python scripts/prof.py --what pipeline --steps 40 2>&1 | tail -2
done
python scripts/prof.py --what rollout_greedy --steps 10 2>&1 | tail -1
export TPL_PAIR=0
P="ncu --set full --clock-control none --import-source on"
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02c_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02c_ncu_fused.log 2>&1
