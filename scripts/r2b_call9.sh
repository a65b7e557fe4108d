set -x
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2b9_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b9_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/overflow_check.sh 2>&1 | tail -4
