# Builds the C-ABI shared library for sm_100a (B200).  `make` == what __graft_entry__.build() runs.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function
SRC := cvxopt_b200/csrc
OBJ := $(SRC)/gemm_dmma.o $(SRC)/chol.o $(SRC)/cone.o $(SRC)/kkt_api.o $(SRC)/blocks_api.o $(SRC)/batch_ipm.o $(SRC)/cone_vec.o $(SRC)/ozaki_syrk.o $(SRC)/nt_scaling.o $(SRC)/kkt_qr.o $(SRC)/kkt_ldl.o
LIB := cvxopt_b200/libcvxopt_b200.so
# CPython extension mirroring cvxopt.misc_solvers over the C ABI (host side of the drop-in boundary)
PYTHON ?= python
PYINC := $(shell $(PYTHON) -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PYEXT := $(shell $(PYTHON) -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
EXT := cvxopt_b200/_misc_solvers$(PYEXT)

all: $(LIB) $(EXT)

$(EXT): $(SRC)/_misc_solvers.c include/cvxopt_b200.h $(LIB)
	gcc -O2 -fPIC -shared -Wall -I$(PYINC) $< -o $@ -Lcvxopt_b200 -lcvxopt_b200 -Wl,-rpath,'$$ORIGIN'

$(SRC)/%.o: $(SRC)/%.cu $(SRC)/common.cuh $(SRC)/cone.cuh $(SRC)/kkt_internal.cuh include/cvxopt_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

# bring-up / timing probe of the experimental int8-slice SYRK (links the library's C++ symbols)
tools/oz_probe: tools/oz_probe.cu $(LIB)
	$(NVCC) $(ARCH) -O2 -std=c++17 -cudart shared $< -o $@ -Lcvxopt_b200 -lcvxopt_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../cvxopt_b200'

clean:
	rm -f $(OBJ) $(LIB) $(EXT) tools/oz_probe
.PHONY: all clean

# roofline-denominator microbenchmarks (run on the GPU box; outputs are committed under profiles/)
tools/int8_peak: tools/int8_peak.cu
	$(NVCC) $(ARCH) -O2 -std=c++17 $< -o $@
tools/fp64_peak: tools/fp64_peak.cu
	$(NVCC) $(ARCH) -O2 -std=c++17 $< -o $@
