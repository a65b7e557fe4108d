set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2b_smoke.log
python scripts/prof.py --what fused --steps 20 > gpurun_out/r2b_prof_fused.log 2>&1
python scripts/prof.py --what fused_distinct --steps 20 > gpurun_out/r2b_prof_fused_distinct.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r2b_fused_distinct python scripts/prof.py --what fused_distinct --steps 3 > gpurun_out/r2b_ncu_fd.log 2>&1
tail -3 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_smoke.log gpurun_out/r2b_prof_*.log
