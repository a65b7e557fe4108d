set -x
python bench.py --steps 20 --warmup 5 --no-dqn > gpurun_out/r2b8_bench.json 2> gpurun_out/r2b8_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b8_bench.err
P="ncu --set full --clock-control none --import-source on"
$P -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r02f_fused python scripts/prof.py --what fused --steps 3 > gpurun_out/r02f_ncu_fused.log 2>&1
