#!/bin/bash
# builds build/variants/<name>.so with extra -D flags: bash tools/build_variant.sh <name> -DHB_NTT16X_MINB8=7 ...
NAME=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" -o build/variants/$NAME.so mpc-protocols_b200/csrc/hbmpc.cu mpc-protocols_b200/csrc/compat_share.cu mpc-protocols_b200/csrc/goldilocks.cu 2>&1 | grep -v "^$" | tail -5
