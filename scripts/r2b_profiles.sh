set -x
bash scripts/r2_ncu_all.sh > gpurun_out/r2b_ncu_all.log 2>&1
tail -8 gpurun_out/r2b_ncu_all.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
python bench.py --steps 5 --warmup 3 --no-dqn > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-dqn > gpurun_out/r02_ncu_bench.log 2>&1
