set -x
timeout 900 python -m pytest tests/test_model.py tests/test_gpu_value.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest.log
python scripts/value_prof.py
python scripts/dqn_bench.py > gpurun_out/r2i_dqn1.json 2> gpurun_out/r2i_dqn1.err; echo "dqn rc=$?"; tail -3 gpurun_out/r2i_dqn1.err; cat gpurun_out/r2i_dqn1.json
OPTIM=4 python scripts/dqn_bench.py > gpurun_out/r2i_dqn4.json 2> gpurun_out/r2i_dqn4.err; cat gpurun_out/r2i_dqn4.json
ITERS=1500 EVERY=250 python scripts/train_demo.py > gpurun_out/r2i_train_demo.txt 2>&1; tail -4 gpurun_out/r2i_train_demo.txt
