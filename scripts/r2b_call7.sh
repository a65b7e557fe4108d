set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "fused_step_observe or rollout_greedy or library_loaded or ragged or edges or long_episodes or deterministic or host or chunk" > gpurun_out/r2b7_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b7_pytest.log
for v in 1 0; do
echo "=== TPL_NO_PDL=$v"
TPL_NO_PDL=$v python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
TPL_NO_PDL=$v python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
TPL_NO_PDL=$v python scripts/prof.py --what pipeline --steps 40 2>&1 | tail -2
done
