# Top-level build: everything is compiled in-tree so that the built libraries
# travel to the GPU box with the snapshot.
#
#   proj_entropy_b200/libgravity_cuda.so   the product: sm_100a kernels + C ABI
#   proj_entropy_b200/libgravity_host.so   host C: scenario reader, Plummer model
#   proj_entropy_b200/gravity_headless     headless C driver (replaces the SDL app)
#   oracle/liboracle.so, oracle/_ref/      TEST-ONLY checkers (see oracle/Makefile)

NVCC      ?= nvcc
CC        ?= gcc
PKG        = proj_entropy_b200
NVCCFLAGS  = -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
             -Xcompiler -fPIC -Iinclude
HOSTFLAGS  = -O2 -fPIC -Wall -Wextra -Iinclude

all: $(PKG)/libgravity_cuda.so $(PKG)/libgravity_host.so $(PKG)/gravity_headless oracle

$(PKG)/libgravity_cuda.so: $(PKG)/csrc/gravity_cuda.cu $(PKG)/csrc/gravity_kernels.cuh include/gravity_cuda.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(PKG)/csrc/gravity_cuda.cu -lnccl

$(PKG)/libgravity_host.so: $(PKG)/host/gravity_host.c include/gravity_host.h
	$(CC) $(HOSTFLAGS) -shared -o $@ $(PKG)/host/gravity_host.c -lm

$(PKG)/gravity_headless: $(PKG)/host/gravity_headless.c $(PKG)/libgravity_host.so $(PKG)/libgravity_cuda.so
	$(CC) $(HOSTFLAGS) -o $@ $(PKG)/host/gravity_headless.c -L$(PKG) -lgravity_host -lgravity_cuda \
	    -Wl,-rpath,'$$ORIGIN' -lm

# after the CUDA library: oracle/_ref/ref_gravity_gpu (the patched reference unit) links against it
oracle: $(PKG)/libgravity_cuda.so $(PKG)/libgravity_host.so
	$(MAKE) -C oracle

tools: tools/fp64_microbench tools/fp64_mix

tools/%: tools/%.cu
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -o $@ $<

clean:
	rm -f $(PKG)/*.so $(PKG)/gravity_headless
	$(MAKE) -C oracle clean

.PHONY: all oracle tools clean
