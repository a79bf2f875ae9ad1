# Builds the product library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC      := recs_b200/csrc
LIB       := recs_b200/librecs_gpu.so
CU        := $(CSRC)/gravity.cu $(CSRC)/streaming.cu $(CSRC)/select.cu $(CSRC)/context.cu
HOSTCPP   := $(wildcard $(CSRC)/host/*.cpp)
OBJ       := $(CU:.cu=.o) $(HOSTCPP:.cpp=.o)
HDR       := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h $(CSRC)/host/*.h) include/recs_gpu.h include/recs_ecs.h

all: $(LIB) oracle examples

$(CSRC)/%.o: $(CSRC)/%.cu $(HDR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(CSRC)/host/%.o: $(CSRC)/host/%.cpp $(HDR)
	g++ -O2 -std=c++17 -fPIC -Wall -Iinclude -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -ldl -lpthread

oracle:
	$(MAKE) -C oracle

# plain-C host on the C ABI alone (no Python): proves the headers are C and the library links standalone
examples: examples/nbody_bench
examples/nbody_bench: examples/nbody_bench.c $(LIB) include/recs_gpu.h include/recs_ecs.h
	$(CC) -O2 -Wall -Iinclude $< -o $@ -Lrecs_b200 -lrecs_gpu -Wl,-rpath,'$$ORIGIN/../recs_b200' -lm

clean:
	rm -f $(OBJ) $(LIB) examples/nbody_bench
	$(MAKE) -C oracle clean

.PHONY: all oracle examples clean
