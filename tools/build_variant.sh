#!/bin/bash
# Builds a variant of libospcuda.so with extra nvcc flags into build/lib_<name>/ (with copies of the two host libraries,
# whose rpath is $ORIGIN) for A/B runs: OSPRAY_B200_LIB_DIR=$PWD/build/lib_<name> python bench.py ...
#   bash tools/build_variant.sh prefetch -DOC_PREFETCH_NEXT
set -e
NAME=$1; shift
OBJ=build/var_$NAME; OUT=build/lib_$NAME
mkdir -p $OBJ $OUT
for f in runtime render fb_ops bvh_build peer_exchange; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
    -c ospray_b200/csrc/cuda/$f.cu -o $OBJ/$f.o &
done
wait
nvcc -shared -o $OUT/libospcuda.so $OBJ/*.o -lcudart
cp ospray_b200/lib/libospray.so ospray_b200/lib/libospray_module_cuda.so $OUT/
echo "built $OUT"
