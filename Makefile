# Builds the product library (CUDA, sm_100a) and the test-side artefacts.
NVCC ?= /usr/local/cuda/bin/nvcc
# the system compiler on purpose: an environment CXX (e.g. /opt/gcc/bin/g++) may link libstdc++ statically, and a second
# copy of libstdc++ in a process that also holds libaffpredict.so / torch breaks iostreams and allocation
CXX = /usr/bin/g++
ARCH = -gencode arch=compute_100a,code=sm_100a
INC = -Iinclude -Iaffaction_b200/csrc -Iaffaction_b200/csrc/host
# -fmad=false: no implicit FMA contraction on the device; the host side gets
# -ffp-contract=off.  Explicit fma() calls stay single instructions on both.
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-variable $(INC)
HOST_SRC = $(wildcard affaction_b200/csrc/host/*.cpp)
KERNEL_HDR = $(wildcard affaction_b200/csrc/kernel/*.h affaction_b200/csrc/kernel/*.cuh) $(wildcard include/*.h) $(wildcard affaction_b200/csrc/host/*.h)

all: affaction_b200/libaffpredict.so affaction_b200/libaffhost.so oracle emu

# the CUDA library: kernels + the C ABI of include/affpredict.h
affaction_b200/libaffpredict.so: affaction_b200/csrc/abi.cu affaction_b200/csrc/abi_common.cpp affaction_b200/csrc/kernel/plan_build.cpp $(KERNEL_HDR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ affaction_b200/csrc/abi.cu affaction_b200/csrc/abi_common.cpp affaction_b200/csrc/kernel/plan_build.cpp 2> build_ptxas.log || (cat build_ptxas.log; false)
	@grep -E "registers|spill|error" build_ptxas.log | head -20

# the host shim (include/affpredict_host.h): pure C++, no CUDA; reaches libaffpredict.so through dlopen on
# the first call that predicts (affaction_b200/csrc/host/gpu_api.cpp).  -Bsymbolic: its C++ symbols bind inside the library (a process
# that has torch loaded otherwise interposes template instantiations of another libstdc++ build: segfault)
affaction_b200/libaffhost.so: $(HOST_SRC) affaction_b200/csrc/abi_common.cpp $(KERNEL_HDR)
	$(CXX) -O2 -std=c++17 -fPIC -shared -ffp-contract=off -Wall -Wno-unused-variable -Wno-maybe-uninitialized $(INC) $(HOST_SRC) affaction_b200/csrc/abi_common.cpp -ldl -pthread -Wl,-Bsymbolic -o $@

oracle:
	$(MAKE) -C oracle

# test-only CPU emulation of the kernel source (never part of the product library)
emu: tests/emu/libaffemu.so
tests/emu/libaffemu.so: tests/emu/warp_emu.cpp tests/emu/flop_count.cpp affaction_b200/csrc/kernel/plan_build.cpp $(KERNEL_HDR)
	$(CXX) -O2 -std=c++17 -fPIC -shared -ffp-contract=off -mfma -Wno-unknown-pragmas $(INC) tests/emu/warp_emu.cpp tests/emu/flop_count.cpp affaction_b200/csrc/kernel/plan_build.cpp -o $@

clean:
	rm -f affaction_b200/libaffpredict.so affaction_b200/libaffhost.so tests/emu/libaffemu.so build_ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle emu clean
