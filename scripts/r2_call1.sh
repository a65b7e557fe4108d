set -x
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; lscpu | grep -i -E "numa|^CPU\(s\)|model name" >> gpurun_out/r2_topo.txt
python -c "import sitecustomize; print(sitecustomize.__file__)" > gpurun_out/r2_hook.txt 2>&1; python - <<'PY' >> gpurun_out/r2_hook.txt 2>&1
import sys, os
for p in sys.path:
    f = os.path.join(p, "sitecustomize.py")
    if os.path.exists(f):
        print("FOUND", f); print(open(f).read()[:6000])
print(os.environ.get("PYTHONPATH"))
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"
python scripts/pcie_bench.py > gpurun_out/r2_pcie_n1.json 2>&1
python scripts/prof.py --what fused --steps 3 > gpurun_out/r2_prof_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_observe_kernel -s 3 -c 1 -o gpurun_out/r2_fused_head python scripts/prof.py --what fused --steps 3 > gpurun_out/r2_ncu_fused.log 2>&1
python scripts/prof.py --what pipeline --steps 3 > gpurun_out/r2_prof_pipe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'afterstates_kernel|step_kernel' -s 6 -c 2 -o gpurun_out/r2_pipe_head python scripts/prof.py --what pipeline --steps 3 > gpurun_out/r2_ncu_pipe.log 2>&1
python scripts/prof.py --what rollout_random --steps 2 > gpurun_out/r2_prof_rr.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/r2_rollout_random_head python scripts/prof.py --what rollout_random --steps 2 > gpurun_out/r2_ncu_rr.log 2>&1
python scripts/prof.py --what rollout_greedy --steps 2 > gpurun_out/r2_prof_rg.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'rollout_kernel<2>|rollout_kernelILi2' -s 2 -c 1 -o gpurun_out/r2_rollout_greedy_head python scripts/prof.py --what rollout_greedy --steps 2 > gpurun_out/r2_ncu_rg.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_b.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_head.csv python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out | tail -30
