set -x
timeout 600 python -m pytest tests/test_gpu_value.py -x -q > gpurun_out/r2d_value.log 2>&1; echo "value rc=$?" | tee -a gpurun_out/r2d_value.log
tail -30 gpurun_out/r2d_value.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_value.py > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
