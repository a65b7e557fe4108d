set -x
timeout 900 python -m pytest tests/test_model.py tests/test_gpu_value.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2g_pytest.log
python scripts/dqn_bench.py > gpurun_out/r2g_dqn.json 2> gpurun_out/r2g_dqn.err; echo "dqn rc=$?"; tail -3 gpurun_out/r2g_dqn.err; cat gpurun_out/r2g_dqn.json
OPTIM=1 python scripts/dqn_bench.py > gpurun_out/r2g_dqn1.json 2> gpurun_out/r2g_dqn1.err; cat gpurun_out/r2g_dqn1.json
python scripts/prof.py --what fused --steps 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:value_rows_kernel -s 2 -c 1 -o gpurun_out/r2g_value python scripts/value_prof.py > gpurun_out/r2g_ncu_value.log 2>&1
