# tuning experiments: rebuild with extra nvcc flags on the GPU box and time the fused step
for flags in "" "-Xptxas --allow-expensive-optimizations=true" "-Xptxas -O4" "-extra-device-vectorization"; do
  TPL_NVCC_EXTRA="$flags" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)
" 2>&1 | tail -1
  grep -A2 "step_observe_kernelILi0ELb1" reinforcement*/csrc/build.log | grep Used
  echo "flags=$flags"; python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
done
