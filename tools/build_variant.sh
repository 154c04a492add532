#!/bin/bash
# A/B builds of the library (development): tools/build_variant.sh NAME "-DMACRO=... ..." -> opencv-simpleslam_b200/libb200slam_NAME.so
# (select with B2S_LIB_PATH).  Only the tensor-core translation units are recompiled with the extra flags.
set -e
NAME=$1; shift
cd "$(dirname "$0")/../opencv-simpleslam_b200/csrc"
make -s -j4
mkdir -p _variants/$NAME
for f in lightglue_tc aliked; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
    --expt-relaxed-constexpr "$@" -c $f.cu -o _variants/$NAME/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libb200slam_$NAME.so api.o lightglue.o geometry.o _variants/$NAME/lightglue_tc.o _variants/$NAME/aliked.o
echo built ../libb200slam_$NAME.so
