set -x
python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
TPL_NVCC_EXTRA="-DTPL_DYNAMIC_TILES=1" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)" > /dev/null 2>&1
echo "dynamic tiles:"
python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or big or compact or deterministic" > gpurun_out/r2k_pytest_dyn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest_dyn.log
