#!/bin/bash
# local pre-flight for a GPU-box call: rebuild the library (the in-tree .so travels with the snapshot), check that it loads
# and exports every declared symbol, then hand the command to gpurun.   usage: tools/grun.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
make -C opencv-simpleslam_b200/csrc -j8 2>&1 | grep -E "error|Error" && exit 1
python -c "import b200slam._lib" 
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
