# Build recipes.  `make lib` = the product (CUDA, sm_100a); `make oracle` = CPU oracle (tests only);
# `make ref` = the reference's own generated path functions compiled into oracle/_ref/ (needs
# /root/reference; outputs are git-ignored but travel to the GPU box).
PKG      := langevin-mcmc_b200
CORE_INC := $(wildcard $(PKG)/csrc/core/*.inc)
DEV_H    := $(wildcard $(PKG)/csrc/core/*.h) $(CORE_INC)
CORE_H   := $(DEV_H) $(wildcard $(PKG)/csrc/host/*.h)
CXX      := $(shell which g++)
CXXFLAGS := -O2 -std=c++17 -fPIC -mfma -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function -pthread
NVCC     := nvcc
NVFLAGS  := -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false \
            -Xcompiler -fPIC,-mfma,-ffp-contract=off,-pthread -Xptxas -v

all: oracle lib

oracle: oracle/liblmc_oracle.so
oracle/liblmc_oracle.so: oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp $(CORE_H)
	$(CXX) $(CXXFLAGS) -shared -o $@ oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp -lz -ldl

# TIMING build of the CPU arm (bench.py cpu_baseline / --impl reference only; never used for parity): the
# reference's own flags (src/Tupfile:17: g++ -march=native -Ofast) and the platform libm.  FAST_ARCH defaults to
# x86-64-v3 so the prebuilt file runs on any AVX2 host; bench.py rebuilds it with -march=native on the box.
FAST_ARCH ?= x86-64-v3
oracle_fast: oracle/liblmc_oracle_fast.so
oracle/liblmc_oracle_fast.so: oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp $(CORE_H)
	$(CXX) -Ofast -march=$(FAST_ARCH) -std=c++17 -fPIC -pthread -DLMC_TIMING_LIBM -w -shared -o $@ oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp -lz -ldl

CUDA_SRC := $(PKG)/csrc/cuda
CUDA_OBJ := $(PKG)/build/lmc_abi.o $(PKG)/build/chain_hess_12.o $(PKG)/build/chain_hess_8.o $(PKG)/build/chain_hess_4.o \
            $(PKG)/build/chain_inst_12.o $(PKG)/build/chain_inst_8.o $(PKG)/build/chain_inst_4.o
lib: $(PKG)/liblmc_b200.so
# the chain kernels see the device headers only; the C ABI translation unit also includes the host side (loader, MLTInit, image io)
$(PKG)/build/lmc_abi.o: $(CORE_H)
$(PKG)/build/%.o: $(CUDA_SRC)/%.cu $(DEV_H) $(wildcard $(CUDA_SRC)/*.cuh) $(wildcard $(CUDA_SRC)/*.h) include/lmc/lmc_abi.h
	@mkdir -p $(PKG)/build
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(PKG)/build/$*.ptxas.log || (cat $(PKG)/build/$*.ptxas.log; false)
$(PKG)/build/host_scene.o: $(PKG)/csrc/host/host_scene.cpp $(CORE_H)
	@mkdir -p $(PKG)/build
	$(CXX) $(CXXFLAGS) -c -o $@ $<
$(PKG)/liblmc_b200.so: $(CUDA_OBJ) $(PKG)/build/host_scene.o
	$(NVCC) -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o $@ $^ -lz -lpthread

ref:
	bash oracle/build_ref.sh

clean:
	rm -rf oracle/liblmc_oracle.so $(PKG)/liblmc_b200.so $(PKG)/build

.PHONY: all oracle oracle_fast lib ref clean
