# usage: bash scripts/r2_multi2.sh N   (weak + strong 8M bench lines at N GPUs)
N=$1
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "weak rc=$?"; tail -3 gpurun_out/r02_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --envs-total 8388608 > gpurun_out/r02_bench_strong8M_n$N.json 2> gpurun_out/r02_bench_strong8M_n$N.err; echo "strong rc=$?"; tail -3 gpurun_out/r02_bench_strong8M_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_n$N.json", "gpurun_out/r02_bench_strong8M_n$N.json"):
    try:
        d = json.load(open(f))
        print(f, d["value"], d["ms_per_step"], d["scaling"], d["collective_us"], "e2e", d["e2e"]["value"], d["e2e"].get("frac_of_measured_pcie_d2h"), "e2e40", d["e2e_40slot"]["value"], d["pcie"])
    except Exception as e:
        print(f, "ERR", e)
PY
