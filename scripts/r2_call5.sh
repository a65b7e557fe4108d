set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2e_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; echo "ref rc=$?"
python scripts/dqn_bench.py > gpurun_out/r2e_dqn.json 2> gpurun_out/r2e_dqn.err; echo "dqn rc=$?"; tail -3 gpurun_out/r2e_dqn.err
