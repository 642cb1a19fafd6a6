#!/bin/bash
# tools/build_exp.sh NAME "-DFLAG ..." : build an experimental variant of the library into build/exp/libcavb200_NAME.so
set -e
NAME=$1; FLAGS=$2
mkdir -p build/exp/$NAME
for f in api hotpath rhok shard host nve track debug; do
  if [ $f = hotpath ] || [ $f = shard ] || [ $f = nve ] || [ $f = rhok ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -diag-suppress 186 $FLAGS -c cav_hoomd_b200/csrc/$f.cu -o build/exp/$NAME/$f.o &
  else
    cp build/obj/$f.o build/exp/$NAME/$f.o
  fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/exp/libcavb200_$NAME.so build/exp/$NAME/*.o -ldl
echo built build/exp/libcavb200_$NAME.so
rm -rf build/exp/$NAME
