TPL_NVCC_EXTRA="-DTPL_VALUE_TRACE=1" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)" > /dev/null 2>&1
python scripts/value_trace.py
