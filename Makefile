# Builds the product library (CUDA, sm_100a) and the test-side artefacts.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH = -gencode arch=compute_100a,code=sm_100a
INC = -Iinclude -Iaffaction_b200/csrc -Iaffaction_b200/csrc/host
# -fmad=false: no implicit FMA contraction on the device; the host side gets
# -ffp-contract=off.  Explicit fma() calls stay single instructions on both.
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall,-Wno-unused-variable $(INC)
HOST_SRC = $(wildcard affaction_b200/csrc/host/*.cpp)
KERNEL_HDR = $(wildcard affaction_b200/csrc/kernel/*.h affaction_b200/csrc/kernel/*.cuh) $(wildcard include/*.h) $(wildcard affaction_b200/csrc/host/*.h)

all: affaction_b200/libaffpredict.so oracle emu

affaction_b200/libaffpredict.so: affaction_b200/csrc/abi.cu affaction_b200/csrc/kernel/plan_build.cpp $(HOST_SRC) $(KERNEL_HDR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared -o $@ affaction_b200/csrc/abi.cu affaction_b200/csrc/kernel/plan_build.cpp $(HOST_SRC) 2> build_ptxas.log || (cat build_ptxas.log; false)
	@grep -E "registers|spill|error" build_ptxas.log | head -20

oracle:
	$(MAKE) -C oracle

# test-only CPU emulation of the kernel source (never part of the product library)
emu: tests/emu/libaffemu.so
tests/emu/libaffemu.so: tests/emu/warp_emu.cpp tests/emu/flop_count.cpp affaction_b200/csrc/kernel/plan_build.cpp $(KERNEL_HDR)
	$(CXX) -O2 -std=c++17 -fPIC -shared -ffp-contract=off -mfma -Wno-unknown-pragmas $(INC) tests/emu/warp_emu.cpp tests/emu/flop_count.cpp affaction_b200/csrc/kernel/plan_build.cpp -o $@

clean:
	rm -f affaction_b200/libaffpredict.so tests/emu/libaffemu.so build_ptxas.log
	$(MAKE) -C oracle clean

.PHONY: all oracle emu clean
