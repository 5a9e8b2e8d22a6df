#!/bin/bash
# usage: tools/build_hess_variant.sh NAME "-DLMC_HESS_MINB=2 ..."   -> langevin-mcmc_b200/liblmc_b200_NAME.so
# (tuning helper: rebuilds only chain_hess_8.cu with extra -D flags and links it with the standard objects)
set -e
cd "$(dirname "$0")/.."
PKG=langevin-mcmc_b200
NAME=$1; shift
mkdir -p $PKG/build/var_$NAME
nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false \
     -Xcompiler -fPIC,-mfma,-ffp-contract=off,-pthread -Xptxas -v $@ \
     -c -o $PKG/build/var_$NAME/chain_hess_8.o $PKG/csrc/cuda/chain_hess_8.cu 2> $PKG/build/var_$NAME/ptxas.log
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o $PKG/liblmc_b200_$NAME.so \
     $PKG/build/lmc_abi.o $PKG/build/chain_hess_4.o $PKG/build/var_$NAME/chain_hess_8.o $PKG/build/chain_hess_12.o \
     $PKG/build/chain_inst_4.o $PKG/build/chain_inst_8.o $PKG/build/chain_inst_12.o $PKG/build/host_scene.o -lz -lpthread
grep -A2 "k_wave_gradILi8ELi2" $PKG/build/var_$NAME/ptxas.log | tail -2
echo built $NAME
