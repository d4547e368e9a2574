# girih-b200: builds the sm_100a library (C ABI), the host library and the mwd_kernel executables.
#
#   make            -> girih_b200/libgirih_cuda.so, girih_b200/libgirih_host_{sp,dp}.so,
#                      build/mwd_kernel (fp32) and build_dp/mwd_kernel (fp64)
#                      (the reference's two-binary convention: `make` / `make dp`, Makefile:12-19)
#   make oracle     -> test tooling under oracle/ (not part of the product)
NVCC      ?= nvcc
HOSTCC    := /usr/bin/gcc
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
             -Xcompiler -fPIC -Xcompiler -Wall
CFLAGS    := -std=gnu99 -O2 -ffp-contract=off -fopenmp -fPIC -Wall -Wno-unknown-pragmas
CSRC      := girih_b200/csrc
HSRC      := girih_b200/host
HOST_C    := $(HSRC)/params.c $(HSRC)/init.c $(HSRC)/steppers.c $(HSRC)/perf.c $(HSRC)/verify.c $(HSRC)/solar.c $(HSRC)/team.c $(HSRC)/pyapi.c
CUDA_DEPS := $(wildcard $(CSRC)/*.cu $(CSRC)/*.cuh $(CSRC)/*.h) include/girih_cuda.h
LIB       := girih_b200/libgirih_cuda.so

.PHONY: all lib host oracle clean
all: lib host

CU_SRC    := $(wildcard $(CSRC)/*.cu)
CU_OBJ    := $(patsubst $(CSRC)/%.cu,$(CSRC)/obj/%.o,$(CU_SRC))

lib: $(LIB)
$(CSRC)/obj/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/girih_cuda.h
	@mkdir -p $(CSRC)/obj
	$(NVCC) $(NVFLAGS) $(PTXASV) -c -o $@ $<
$(LIB): $(CU_OBJ)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -shared -o $@ $(CU_OBJ) -ldl

host: build/mwd_kernel build_dp/mwd_kernel girih_b200/libgirih_host_sp.so girih_b200/libgirih_host_dp.so

build/mwd_kernel: $(HOST_C) $(HSRC)/driver.c $(HSRC)/girih_host.h $(LIB)
	mkdir -p build
	$(HOSTCC) $(CFLAGS) -DDP=0 -o $@ $(HOST_C) $(HSRC)/driver.c -Lgirih_b200 -lgirih_cuda -Wl,-rpath,'$$ORIGIN/../girih_b200' -lpthread -lm

build_dp/mwd_kernel: $(HOST_C) $(HSRC)/driver.c $(HSRC)/girih_host.h $(LIB)
	mkdir -p build_dp
	$(HOSTCC) $(CFLAGS) -DDP=1 -o $@ $(HOST_C) $(HSRC)/driver.c -Lgirih_b200 -lgirih_cuda -Wl,-rpath,'$$ORIGIN/../girih_b200' -lpthread -lm

girih_b200/libgirih_host_sp.so: $(HOST_C) $(HSRC)/girih_host.h $(LIB)
	$(HOSTCC) $(CFLAGS) -DDP=0 -shared -o $@ $(HOST_C) -Lgirih_b200 -lgirih_cuda -Wl,-rpath,'$$ORIGIN' -lpthread -lm

girih_b200/libgirih_host_dp.so: $(HOST_C) $(HSRC)/girih_host.h $(LIB)
	$(HOSTCC) $(CFLAGS) -DDP=1 -shared -o $@ $(HOST_C) -Lgirih_b200 -lgirih_cuda -Wl,-rpath,'$$ORIGIN' -lpthread -lm

oracle:
	$(MAKE) -C oracle oracle
	if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref -j8; fi

clean:
	rm -rf build build_dp girih_b200/*.so girih_b200/csrc/obj
