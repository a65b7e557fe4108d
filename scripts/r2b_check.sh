set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_step_observe or rollout_greedy or library_loaded or ragged or edges or long_episodes or deterministic" > gpurun_out/r2b_check_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_check_pytest.log
python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
python scripts/prof.py --what pipeline --steps 40 2>&1 | tail -2
python scripts/prof.py --what rollout_greedy --steps 10 2>&1 | tail -1
P="ncu --set full --clock-control none --import-source on"
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02e_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02e_ncu_fused.log 2>&1
