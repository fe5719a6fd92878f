# Builds libmmpgo.so (the C-ABI CUDA library) in-tree for sm_100a and the C oracle helpers.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp,-O3 -Iinclude
SRC := dpgo_b200/csrc/mmpgo_kernels.cu dpgo_b200/csrc/mmpgo_tsolve.cu dpgo_b200/csrc/mmpgo_setup.cu dpgo_b200/csrc/mmpgo_factor.cu dpgo_b200/csrc/mmpgo_mfsolve.cu dpgo_b200/csrc/mmpgo_driver.cu dpgo_b200/csrc/mmpgo_nccl.cu dpgo_b200/csrc/mmpgo_capi.cu
OBJ := $(SRC:.cu=.o)
LIB := dpgo_b200/libmmpgo.so

HOSTBIN := host/dist_pgo

all: $(LIB) $(HOSTBIN)

# the reference-language host: dist_pgo CLI + DPGOHash/DPGOStar classes over the C ABI
$(HOSTBIN): host/src/dist_pgo.cpp host/include/mmpgo_host/DPGO.h include/mmpgo.h $(LIB)
	g++ -O2 -std=c++17 -fopenmp -Ihost/include host/src/dist_pgo.cpp -o $@ -Ldpgo_b200 -lmmpgo -Wl,-rpath,'$$ORIGIN/../dpgo_b200'

%.o: %.cu dpgo_b200/csrc/mmpgo_kernels.cuh dpgo_b200/csrc/mmpgo_driver.cuh dpgo_b200/csrc/so3_project.cuh dpgo_b200/csrc/mmpgo_mf.cuh include/mmpgo.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -Xcompiler -fopenmp -lgomp -ldl -cudart shared

clean:
	rm -f $(OBJ) $(LIB) $(HOSTBIN)
.PHONY: all clean
