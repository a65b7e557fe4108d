for mb in 4 5 6; do
  TPL_NVCC_EXTRA="-DTPL_AS_MINBLOCKS=$mb" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)
" 
  grep -A2 "step_observe_kernelILi0" reinforcement*/csrc/build.log | grep Used
  echo "MINBLOCKS=$mb"; python scripts/prof.py --what fused --steps 30 2>&1 | tail -1
done
