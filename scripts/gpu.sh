#!/bin/bash
# Rebuild every native piece locally (nvcc cross-compiles), then run a command on a B200 box.
# usage: scripts/gpu.sh <gpurun-timeout-seconds> '<command>' [extra gpurun flags]
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" >/dev/null
T=$1; shift
CMD=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
