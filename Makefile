# Builds the product library (CUDA, sm_100a only) and the test-only oracle (plain C).
#   make            -> zero-shot-aac_b200/lib/libzsaac_b200.so  +  oracle/liboracle.so
#   make lib        -> the CUDA library only
#   make oracle     -> the C restatement used by tests / bench cpu_baseline only
NVCC      ?= nvcc
CC        ?= gcc
PKG       := zero-shot-aac_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/lib/libzsaac_b200.so
ORACLE    := oracle/liboracle.so
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -Wall -shared -cudart static

all: lib oracle

lib: $(LIB)

$(LIB): $(CSRC)/zsaac_api.cu $(CSRC)/simtopk_kernel.cuh $(CSRC)/aux_kernels.cuh $(CSRC)/ptx_sm100.cuh include/zsaac.h
	@mkdir -p $(PKG)/lib
	$(NVCC) $(NVCCFLAGS) $(EXTRA_NVCCFLAGS) -o $@ $(CSRC)/zsaac_api.cu

oracle: $(ORACLE)

$(ORACLE): oracle/oracle_topk.c
	$(CC) -O3 -march=x86-64-v2 -fopenmp -fPIC -shared -o $@ $< -lm

clean:
	rm -f $(LIB) $(ORACLE)

.PHONY: all lib oracle clean
