# Builds the product library: hand-written CUDA for sm_100a behind the C ABI of include/zsaac.h.
#   make / make lib  -> zero-shot-aac_b200/lib/libzsaac_b200.so   (cross-compiles without a GPU)
# The oracle (oracle/oracle.py) is pure Python and needs no build step.
NVCC      ?= nvcc
PKG       := zero-shot-aac_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/lib/libzsaac_b200.so
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -Wall -shared -cudart static

all: lib

lib: $(LIB)

$(LIB): $(CSRC)/zsaac_api.cu $(CSRC)/memproj_kernel.cuh $(CSRC)/memproj_tc_kernels.cuh $(CSRC)/simtopk_kernel.cuh $(CSRC)/aux_kernels.cuh $(CSRC)/exact_f32_kernels.cuh $(CSRC)/ptx_sm100.cuh include/zsaac.h
	@mkdir -p $(PKG)/lib
	$(NVCC) $(NVCCFLAGS) $(EXTRA_NVCCFLAGS) -o $@ $(CSRC)/zsaac_api.cu

clean:
	rm -f $(LIB)

.PHONY: all lib clean
