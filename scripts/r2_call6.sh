set -x
python scripts/dqn_profile.py 1 > gpurun_out/r2f_dqn_profile_vk.txt 2>&1; echo rc=$?
python scripts/prof.py --what fused --steps 40 > gpurun_out/r2f_prof.log 2>&1; python scripts/prof.py --what fused_distinct --steps 40 >> gpurun_out/r2f_prof.log 2>&1; cat gpurun_out/r2f_prof.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2f_bench.err
