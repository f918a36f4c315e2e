# Top-level build: libsrcnn_b200.so (the product: CUDA kernels + C ABI), bin/srcnn (the drop-in CLI),
# and the oracle checkers (test infrastructure).  sm_100a only; no other arch, no fallback.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := /usr/bin/g++
HOSTCC := /usr/bin/gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math \
           --fmad=true -Xptxas -v -Wno-deprecated-gpu-targets
CSRC := srcnn_cpp_b200/csrc
OBJ := build/obj
LIB := srcnn_cpp_b200/libsrcnn_b200.so
WEIGHTS := $(abspath srcnn_cpp_b200/data/srcnn_weights.bin)
CU := api mgpu jpeg_stream color_bicubic color_bicubic_int srcnn_fp32 srcnn_tc2 fraw_resize
OBJS := $(addprefix $(OBJ)/,$(addsuffix .o,$(CU))) $(OBJ)/weights_blob.o $(OBJ)/libsrcnn.o

all: $(LIB) bin/srcnn oracle

$(OBJ)/%.o: $(CSRC)/%.cu $(CSRC)/common.h $(CSRC)/color_bicubic.h $(CSRC)/weights.h include/srcnn_b200.h
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)
	@grep -E "error|warning|spill|registers" $(OBJ)/$*.ptxas.log | grep -v "0 bytes spill" | head -40 || true

$(OBJ)/weights_blob.o: $(CSRC)/weights_blob.c $(WEIGHTS)
	@mkdir -p $(OBJ)
	$(HOSTCC) -c -fPIC -DSRCNN_WEIGHTS_BIN='"$(WEIGHTS)"' $< -o $@

$(OBJ)/libsrcnn.o: srcnn_cpp_b200/cli/libsrcnn.cpp include/libsrcnn.h include/srcnn_b200.h
	@mkdir -p $(OBJ)
	$(HOSTCXX) -std=c++17 -O2 -fPIC -fvisibility=hidden -c $< -o $@

bin/srcnn: srcnn_cpp_b200/cli/srcnn_main.cpp srcnn_cpp_b200/cli/image_io.cpp srcnn_cpp_b200/cli/jpeg_io.cpp srcnn_cpp_b200/cli/image_io.h $(LIB)
	@mkdir -p bin
	$(HOSTCXX) -std=c++17 -O2 -I/usr/local/cuda/include srcnn_cpp_b200/cli/srcnn_main.cpp srcnn_cpp_b200/cli/image_io.cpp \
	    srcnn_cpp_b200/cli/jpeg_io.cpp -o $@ -Lsrcnn_cpp_b200 -lsrcnn_b200 -L/usr/local/cuda/lib64 -lnvjpeg_static -lculibos -lcudart_static \
	    -ldl -lrt -lz -lpthread -Wl,-rpath,'$$ORIGIN/../srcnn_cpp_b200'

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -Xcompiler -fPIC -o $@ $(OBJS) -lnvjpeg_static -lculibos -lcuda

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) bin
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
