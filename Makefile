# Builds the product: lulesh_b200/lib/liblulesh_b200.so (sm_100a kernels + C ABI +
# host Domain) and lulesh_b200/bin/lulesh_b200 (the drop-in driver).  In-tree so
# the artefacts travel to the GPU box with the snapshot.
NVCC    ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(shell command -v /usr/bin/g++ || echo g++)
ARCH    := -gencode arch=compute_100a,code=sm_100a
KDEFS   ?=
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin $(HOSTCXX) $(KDEFS)
CXXFLAGS := -O2 -std=c++17 -fPIC -ffp-contract=off -Wall
CSRC := lulesh_b200/csrc
OBJ  := build
LIB  := lulesh_b200/lib/liblulesh_b200.so
BIN  := lulesh_b200/bin/lulesh_b200

all: $(LIB) $(BIN)

$(OBJ)/kernels.o: $(CSRC)/kernels.cu $(CSRC)/kernels.cuh include/lulesh_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/kernels.ptxas.log || (cat $(OBJ)/kernels.ptxas.log; false)

$(OBJ)/api.o: $(CSRC)/api.cu $(CSRC)/kernels.cuh $(CSRC)/setup.cuh $(CSRC)/host/domain.h include/lulesh_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/setup.o: $(CSRC)/setup.cu $(CSRC)/setup.cuh $(CSRC)/kernels.cuh include/lulesh_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(OBJ)/%.o: $(CSRC)/host/%.cc $(CSRC)/host/domain.h include/lulesh_b200.h include/lulesh_host.h
	@mkdir -p $(OBJ)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(OBJ)/kernels.o $(OBJ)/api.o $(OBJ)/setup.o $(OBJ)/domain.o $(OBJ)/host_capi.o $(OBJ)/driver.o $(OBJ)/vizdump.o
	@mkdir -p lulesh_b200/lib
	$(NVCC) $(ARCH) -shared -cudart static -ccbin $(HOSTCXX) -o $@ $^ -ldl -lpthread

$(BIN): $(OBJ)/main.o $(LIB)
	@mkdir -p lulesh_b200/bin
	$(HOSTCXX) -o $@ $(OBJ)/main.o -Llulesh_b200/lib -llulesh_b200 -Wl,-rpath,'$$ORIGIN/../lib'

clean:
	rm -rf $(OBJ) $(LIB) $(BIN)

.PHONY: all clean
