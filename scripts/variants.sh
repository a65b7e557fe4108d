# tuning experiments: rebuild with extra nvcc flags on the GPU box and time the fused step (40-slot compact and distinct forms)
# usage: bash scripts/variants.sh "<flags 1>" "<flags 2>" ...      ("" = the shipped build)
for flags in "$@"; do
  TPL_NVCC_EXTRA="$flags" python -c "
import importlib,sys
sys.path.insert(0,'.')
b=importlib.import_module('reinforcement-learning-for-playing-tetris-with-prescribed-initial-configuration-and-limited-moves_b200.build')
b.build(force=True)
" 2>&1 | tail -1
  echo "=== flags=[$flags]"
  grep -E "step_observe_kernelILi(0ELb1|4)" -A2 lib/build.log | grep Used
  python scripts/prof.py --what fused --steps 40 2>&1 | tail -1
  python scripts/prof.py --what fused_distinct --steps 40 2>&1 | tail -1
done
