# Builds the sm_100a C-ABI library in-tree (the .so travels to the GPU box with the snapshot).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(EXTRA) -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC := signaltrain_b200/csrc
OBJ := $(CSRC)/st_api.o $(CSRC)/st_frontend.o $(CSRC)/st_gemm_simt.o $(CSRC)/st_gemm_tc.o $(CSRC)/st_ae.o $(CSRC)/st_ae_mma.o $(CSRC)/st_ae_tm.o $(CSRC)/st_loss_opt.o $(CSRC)/st_data.o
LIB := signaltrain_b200/lib/libsignaltrain_b200.so

all: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/st_common.cuh $(CSRC)/st_tc_prims.cuh include/signaltrain_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(CSRC)/$*.ptxas.log || (cat $(CSRC)/$*.ptxas.log; false)

$(LIB): $(OBJ)
	mkdir -p signaltrain_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

clean:
	rm -f $(OBJ) $(LIB) $(CSRC)/*.ptxas.log

.PHONY: all clean

# stand-alone kernel harness (test infrastructure; run under gpurun)
HARNESS := tests/native/ae_tm_harness
$(HARNESS): tests/native/ae_tm_harness.cu $(CSRC)/st_ae_tm.o
	$(NVCC) -O2 -std=c++17 $(ARCH) -o $@ $^ -lcudart
harness: $(HARNESS)
