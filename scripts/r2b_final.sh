set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/r2_ncu_all.sh > gpurun_out/r2b_final_ncu_all.log 2>&1; tail -8 gpurun_out/r2b_final_ncu_all.log
