set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pair_tile or fused_step_full_size or library_loaded" > gpurun_out/r2b1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2b1_pytest.log
for p in 0 1; do
  echo "=== TPL_PAIR=$p"
  TPL_PAIR=$p python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
  TPL_PAIR=$p python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
  TPL_PAIR=$p python scripts/prof.py --what pipeline --steps 40 2>&1 | tail -2
done
