# Builds the C-ABI shared library for sm_100a (B200).  `make` == what __graft_entry__.build() runs.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
SRC := cvxopt_b200/csrc
OBJ := $(SRC)/gemm_dmma.o $(SRC)/chol.o $(SRC)/cone.o $(SRC)/kkt_api.o $(SRC)/blocks_api.o $(SRC)/batch_ipm.o $(SRC)/cone_vec.o $(SRC)/ozaki_syrk.o
LIB := cvxopt_b200/libcvxopt_b200.so

all: $(LIB)

$(SRC)/%.o: $(SRC)/%.cu $(SRC)/common.cuh $(SRC)/cone.cuh include/cvxopt_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

# bring-up / timing probe of the experimental int8-slice SYRK (links the library's C++ symbols)
tools/oz_probe: tools/oz_probe.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -cudart shared $< -o $@ -Lcvxopt_b200 -lcvxopt_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../cvxopt_b200'

clean:
	rm -f $(OBJ) $(LIB) tools/oz_probe
.PHONY: all clean
