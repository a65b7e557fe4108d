set -x
python scripts/value_prof.py > gpurun_out/r2_prof_value.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:value_rows_kernel -s 2 -c 1 -o gpurun_out/r02_value python scripts/value_prof.py > gpurun_out/r02_ncu_value.log 2>&1
cat gpurun_out/r2_prof_value.log
python scripts/dqn_bench.py > gpurun_out/r2l_dqn1.json 2> gpurun_out/r2l_dqn1.err; cat gpurun_out/r2l_dqn1.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
