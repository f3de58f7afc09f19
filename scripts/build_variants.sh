#!/bin/bash
# Build A/B variants of libiwae_b200.so with extra -D flags: scripts/build_variants.sh name1="-DX=1" name2="-DY=2 -DZ" ...
# -> iwae_b200/lib/variants/libiwae_<name>.so (git-ignored; travels to the GPU box).  Select with IWAE_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../iwae_b200/csrc"
mkdir -p ../lib/variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  rm -rf build_var && mkdir -p build_var
  for f in abi kernels_f32 eval_tc encoder_tc comm tools; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c $f.cu -o build_var/$f.o &
  done
  wait
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libiwae_$name.so build_var/abi.o build_var/kernels_f32.o build_var/eval_tc.o build_var/encoder_tc.o build_var/comm.o build_var/tools.o -ldl
  rm -rf build_var
  echo "built variant $name ($flags)"
done
