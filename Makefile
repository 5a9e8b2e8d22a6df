# Build recipes.  `make lib` = the product (CUDA, sm_100a); `make oracle` = CPU oracle (tests only);
# `make ref` = the reference's own generated path functions compiled into oracle/_ref/ (needs
# /root/reference; outputs are git-ignored but travel to the GPU box).
PKG      := langevin-mcmc_b200
CORE_H   := $(wildcard $(PKG)/csrc/core/*.h) $(wildcard $(PKG)/csrc/host/*.h)
CXX      := $(shell which g++)
CXXFLAGS := -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -Wall -Wno-unused-function -pthread
NVCC     := nvcc
NVFLAGS  := -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false \
            -Xcompiler -fPIC,-ffp-contract=off,-pthread -Xptxas -v

all: oracle lib

oracle: oracle/liblmc_oracle.so
oracle/liblmc_oracle.so: oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp $(CORE_H)
	$(CXX) $(CXXFLAGS) -shared -o $@ oracle/oracle_api.cpp $(PKG)/csrc/host/host_scene.cpp -lz

lib: $(PKG)/liblmc_b200.so
$(PKG)/liblmc_b200.so: $(PKG)/csrc/cuda/lmc_abi.cu $(PKG)/csrc/host/host_scene.cpp $(CORE_H) include/lmc/lmc_abi.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(PKG)/csrc/cuda/lmc_abi.cu $(PKG)/csrc/host/host_scene.cpp -lz 2> $(PKG)/ptxas.log || (cat $(PKG)/ptxas.log; false)

ref:
	bash oracle/build_ref.sh

clean:
	rm -f oracle/liblmc_oracle.so $(PKG)/liblmc_b200.so

.PHONY: all oracle lib ref clean
