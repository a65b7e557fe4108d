set -x
timeout 600 python -m pytest tests/test_gpu_value.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2j_pytest.log
python scripts/value_prof.py
TPL_NVCC_EXTRA="-DTPL_VALUE_A_TMEM=0" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)" > /dev/null 2>&1
echo "A in shared memory:"; python scripts/value_prof.py
