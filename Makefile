# Builds libd2gpu.so (CUDA kernels + C ABI, sm_100a only) in-tree under dashing2_b200/.
# One object per translation unit (make -j builds them in parallel); objects and ptxas logs go to build/.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX_HOST ?= /usr/bin/g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -ccbin $(CXX_HOST) --fmad=false -Xptxas -v
CSRC := dashing2_b200/csrc
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/d2gpu.h
UNITS := api_core api_comm api_sketch api_stream api_weighted api_cmp api_lsh
OBJS := $(patsubst %,build/%.o,$(UNITS)) build/pack_host.o

all: dashing2_b200/libd2gpu.so dashing2_b200/bin/dashing2-gpu

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

build/pack_host.o: $(CSRC)/host/pack_host.cpp $(CSRC)/host/pack_host.h
	@mkdir -p build
	$(CXX_HOST) -O3 -std=c++17 -fPIC -Wall -c -o $@ $<

dashing2_b200/libd2gpu.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(CXX_HOST) -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# drop-in front-end for `dashing2 sketch|cmp` (host C++ only; all numerics are in libd2gpu)
dashing2_b200/bin/dashing2-gpu: $(CSRC)/host/d2_main.cpp include/d2gpu.h dashing2_b200/libd2gpu.so
	@mkdir -p dashing2_b200/bin
	$(CXX_HOST) -O2 -std=c++17 -Wall -o $@ $(CSRC)/host/d2_main.cpp -Ldashing2_b200 -ld2gpu -lz -lpthread -Wl,-rpath,'$$ORIGIN/..'

clean:
	rm -rf build dashing2_b200/libd2gpu.so
.PHONY: all clean
