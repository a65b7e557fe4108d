# usage: bash scripts/r2_multi.sh N   (N GPUs of one box)
N=$1
set -x
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -q -k non_current > gpurun_out/r02_multi_dev_test_n$N.log 2>&1; tail -2 gpurun_out/r02_multi_dev_test_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "weak rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --envs-total 8388608 > gpurun_out/r02_bench_strong8M_n$N.json 2> gpurun_out/r02_bench_strong8M_n$N.err; echo "strong rc=$?"; tail -3 gpurun_out/r02_bench_strong8M_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/pcie_bench.py > gpurun_out/r02_pcie_n$N.json 2>&1; cat gpurun_out/r02_pcie_n$N.json
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n$N.json", "gpurun_out/r02_bench_strong8M_n$N.json"):
    try:
        d = json.load(open(f))
        print(f, d["value"], d["ms_per_step"], d["scaling"], d["collective_us"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_measured_pcie_d2h"), "e2e40", d["e2e_40slot"]["value"], d["pcie"])
    except Exception as e:
        print(f, "ERR", e)
PY
