# Builds libxspect_b200.so (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared
LIB := xspect2_b200/libxspect_b200.so
SRC := xspect2_b200/csrc/xs_lib.cu xspect2_b200/csrc/xs_fastx.cpp xspect2_b200/csrc/xs_result.cpp
HDR := xspect2_b200/csrc/xs_kernels.cuh xspect2_b200/csrc/xs_device.cuh include/xspect_b200.h

all: $(LIB) oracle hostcheck

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -o $@ $(SRC)

oracle:
	$(MAKE) -C oracle libxs_oracle.so

hostcheck: tests/native/libxs_hostcheck.so
tests/native/libxs_hostcheck.so: tests/native/xs_hostcheck.cu xspect2_b200/csrc/xs_device.cuh
	$(NVCC) -O2 -std=c++17 -Xcompiler -fPIC -shared -o $@ $<

ptxas:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -o /tmp/xs_ptxas.so $(SRC)

clean:
	rm -f $(LIB) tests/native/libxs_hostcheck.so oracle/libxs_oracle.so

.PHONY: all oracle hostcheck ptxas clean
