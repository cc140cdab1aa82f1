# Builds the product shared library (CUDA kernels + C ABI + host-side mirror) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := eagle-mpc_b200
LIB := $(PKG)/lib/libempc_b200.so
CUFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall
HOST_SRC := $(PKG)/host/params.cpp $(PKG)/host/urdf.cpp $(PKG)/host/trajectory.cpp $(PKG)/host/sbfddp.cpp $(PKG)/host/mpc.cpp $(PKG)/host/host_capi.cpp
HOST_OBJ := $(HOST_SRC:.cpp=.o)
CU_OBJ := $(PKG)/csrc/solver.o

all: $(LIB) oracle

$(PKG)/csrc/solver.o: $(PKG)/csrc/solver.cu $(wildcard $(PKG)/csrc/*.cuh) include/empc_b200.h
	$(NVCC) $(CUFLAGS) -c -o $@ $<

$(PKG)/host/%.o: $(PKG)/host/%.cpp $(PKG)/host/eagle_mpc.hpp $(PKG)/host/mpc.hpp include/empc_b200.h
	g++ $(CXXFLAGS) -c -o $@ $<

$(LIB): $(CU_OBJ) $(HOST_OBJ)
	mkdir -p $(PKG)/lib
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJ) $(HOST_OBJ) -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(PKG)/csrc/*.o $(PKG)/host/*.o $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
