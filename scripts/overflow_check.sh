# one-off validation of the deferred-slot queue's overflow path: rebuild with a 72-entry queue (the smallest allowed), run the parity tests
TPL_NVCC_EXTRA="-DTPL_WQ_ITEMS=72" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)
"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused or afterstates or golden" 2>&1 | tail -3
python scripts/prof.py --what fused --steps 20 2>&1 | tail -1
