#!/bin/bash
# Builds A/B variants of the library with extra -D flags: bash tools/build_variants.sh name1 "-DX=1" name2 "-DY=2 -DZ=3" ...
# -> build/ab/libape_<name>.so (git-ignored; they travel to the GPU box).  Compare with tools/lib_ab.sh.
cd "$(dirname "$0")/.."
mkdir -p build/ab
SRCS=$(ls arm_pose_estimation_b200/csrc/*.cu)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude $flags -o build/ab/libape_$name.so $SRCS 2>&1 | grep -i "error" &
done
wait
ls -la build/ab
