# Builds the product shared library (CUDA kernels + C ABI + host-side mirror) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := eagle-mpc_b200
LIB := $(PKG)/lib/libempc_b200.so
# host-side mirror of the reference's C++ surfaces (YAML / URDF / Trajectory / MPC controllers): no CUDA code, binds the
# C ABI of $(LIB) at run time (host/cuda_abi.cpp), so problem construction never loads the CUDA library
HOSTLIB := $(PKG)/lib/libempc_host.so
CUFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall
HOST_SRC := $(PKG)/host/params.cpp $(PKG)/host/urdf.cpp $(PKG)/host/trajectory.cpp $(PKG)/host/sbfddp.cpp $(PKG)/host/mpc.cpp $(PKG)/host/host_capi.cpp $(PKG)/host/cuda_abi.cpp
HOST_OBJ := $(HOST_SRC:.cpp=.o)
CU_OBJ := $(PKG)/csrc/solver.o

PYEXT := $(PKG)/python/eagle_mpc/_eagle_mpc$(shell python3 -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC := $(shell python3 -c "import pybind11, sysconfig; print('-I' + pybind11.get_include() + ' -I' + sysconfig.get_paths()['include'])")

all: $(LIB) $(HOSTLIB) $(PYEXT) oracle

$(PKG)/csrc/solver.o: $(PKG)/csrc/solver.cu $(wildcard $(PKG)/csrc/*.cuh) include/empc_b200.h
	$(NVCC) $(CUFLAGS) -c -o $@ $<

$(PKG)/host/%.o: $(PKG)/host/%.cpp $(PKG)/host/eagle_mpc.hpp $(PKG)/host/mpc.hpp $(PKG)/host/cuda_abi.hpp include/empc_b200.h
	g++ $(CXXFLAGS) -c -o $@ $<

$(LIB): $(CU_OBJ)
	mkdir -p $(PKG)/lib
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJ)

$(HOSTLIB): $(HOST_OBJ)
	mkdir -p $(PKG)/lib
	g++ -shared -o $@ $(HOST_OBJ) -ldl

# Python front-end with the reference's names (pybind11 over the host mirror; binds the CUDA library at run time)
$(PYEXT): $(PKG)/host/pybind.cpp $(HOST_OBJ)
	g++ $(CXXFLAGS) -shared -fvisibility=hidden $(PYINC) -o $@ $(PKG)/host/pybind.cpp $(filter-out $(PKG)/host/host_capi.o,$(HOST_OBJ)) -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(PKG)/csrc/*.o $(PKG)/host/*.o $(LIB) $(HOSTLIB) $(PYEXT)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
