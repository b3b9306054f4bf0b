# Builds libb2resample.so (CUDA, sm_100a only) and the b2resample CLI in-tree.
# Usage: make -j8            (nvcc cross-compiles without a GPU)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 -lineinfo $(ARCH) --expt-relaxed-constexpr -Xcompiler -fPIC $(NVCC_EXTRA)
SRC       := vkresample_b200/csrc
OUT       := vkresample_b200/lib
OBJ       := build/obj
CU_SRCS   := b2r_api.cu b2r_static_r2c.cu b2r_static_cols.cu b2r_dynamic.cu b2r_sharpen.cu
# the two slowest translation units are compiled in parts (same source, one -D per part) to use more cores
C2R_PARTS := 0 1 2 3
DYN_CCS   := 1 2 4 8
OBJS      := $(addprefix $(OBJ)/,$(CU_SRCS:.cu=.o)) $(OBJ)/b2r_plan.o $(OBJ)/b2r_jit.o \
             $(foreach p,$(C2R_PARTS),$(OBJ)/b2r_static_c2r_$(p).o) $(foreach c,$(DYN_CCS),$(OBJ)/b2r_dynamic_cols_$(c).o)
HDRS      := $(wildcard $(SRC)/*.cuh $(SRC)/*.h) include/b2resample.h

all: $(OUT)/libb2resample.so $(OUT)/b2resample

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

$(OBJ)/b2r_static_c2r_%.o: $(SRC)/b2r_static_c2r.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DB2R_C2R_PART=$* -Xptxas -v -c $< -o $@ 2> $(OBJ)/b2r_static_c2r_$*.ptxas.log || (cat $(OBJ)/b2r_static_c2r_$*.ptxas.log; exit 1)

$(OBJ)/b2r_dynamic_cols_%.o: $(SRC)/b2r_dynamic_cols.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -DB2R_DYN_CC=$* -Xptxas -v -c $< -o $@ 2> $(OBJ)/b2r_dynamic_cols_$*.ptxas.log || (cat $(OBJ)/b2r_dynamic_cols_$*.ptxas.log; exit 1)

$(OBJ)/b2r_plan.o: $(SRC)/b2r_plan.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(OBJ)/b2r_jit.o: $(SRC)/b2r_jit.cpp $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(OUT)/libb2resample.so: $(OBJS)
	@mkdir -p $(OUT)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

$(OUT)/b2resample: $(SRC)/cli.cpp $(OUT)/libb2resample.so include/b2resample.h
	$(CXX) -std=c++17 -O3 -Iinclude $< -o $@ -L$(OUT) -lb2resample -lz -lpthread -Wl,-rpath,'$$ORIGIN'

emu:
	tests/emu/build.sh

clean:
	rm -rf build $(OUT)/libb2resample.so $(OUT)/b2resample

.PHONY: all emu clean
