# Build the product library (sm_100a CUDA + C host layer) and the test oracle.
#   make            -> schroedinger_b200/libschro_b200.so + oracle/liboracle.so (+ oracle/_ref when the reference is present)
# nvcc cross-compiles without a GPU.
NVCC      ?= nvcc
CC        ?= gcc
CUDA_HOME ?= /usr/local/cuda
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) $(EXTRA_NVFLAGS) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Ischroedinger_b200/csrc
CFLAGS    := -std=gnu99 -O2 -Wall -fPIC -Iinclude -I$(CUDA_HOME)/include

CU_SRCS   := $(wildcard schroedinger_b200/csrc/*.cu)
CU_OBJS   := $(patsubst schroedinger_b200/csrc/%.cu,build/%.o,$(CU_SRCS))
HOST_SRCS := $(wildcard schroedinger_b200/host/*.c)
HOST_OBJS := $(patsubst schroedinger_b200/host/%.c,build/host_%.o,$(HOST_SRCS))
ORACLE_SRCS := $(wildcard oracle/oracle_*.c)

all: schroedinger_b200/libschro_b200.so bench_native/libsb2_e2e_driver.so oracle/liboracle.so ref

build/%.o: schroedinger_b200/csrc/%.cu $(wildcard schroedinger_b200/csrc/*.cuh) include/schro_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

build/host_%.o: schroedinger_b200/host/%.c $(wildcard include/*.h) $(wildcard schroedinger_b200/host/*.h)
	@mkdir -p build
	$(CC) $(CFLAGS) -c $< -o $@

schroedinger_b200/libschro_b200.so: $(CU_OBJS) $(HOST_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

# pthread driver of bench.py's end-to-end leg: plain C against the drop-in API
bench_native/libsb2_e2e_driver.so: bench_native/e2e_driver.c include/schro_b200_compat.h schroedinger_b200/libschro_b200.so
	$(CC) -std=gnu99 -O2 -Wall -fPIC -shared -Iinclude -o $@ bench_native/e2e_driver.c \
	    -Lschroedinger_b200 -lschro_b200 -lpthread -Wl,-rpath,'$$ORIGIN/../schroedinger_b200'

oracle/liboracle.so: $(ORACLE_SRCS) oracle/oracle.h $(wildcard oracle/*.inc)
	$(CC) -std=gnu99 -O2 -Wall -fPIC -fwrapv -ffp-contract=off -shared -o $@ $(ORACLE_SRCS)

ref:
	bash oracle/build_ref.sh

clean:
	rm -rf build schroedinger_b200/libschro_b200.so bench_native/libsb2_e2e_driver.so oracle/liboracle.so oracle/_ref

.PHONY: all ref clean
