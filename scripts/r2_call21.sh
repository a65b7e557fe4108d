set -x
python scripts/value_prof.py > gpurun_out/r2_prof_value.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:value_rows_kernel -s 2 -c 1 -o gpurun_out/r02_value python scripts/value_prof.py > gpurun_out/r02_ncu_value.log 2>&1
cat gpurun_out/r2_prof_value.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2m_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 5 --warmup 3 --no-dqn > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-dqn > gpurun_out/r02_ncu_bench.log 2>&1
