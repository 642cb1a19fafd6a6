# Top-level build: libcavb200.so (the product), the CPU oracle (test infrastructure) and the
# plugin glue built against hoomd_shim (test build; a real deployment uses plugin/CMakeLists.txt).
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -diag-suppress 186
CSRC      := cav_hoomd_b200/csrc
LIBDIR    := cav_hoomd_b200/lib
OBJDIR    := build/obj
SRCS      := $(CSRC)/api.cu $(CSRC)/hotpath.cu $(CSRC)/rhok.cu $(CSRC)/shard.cu $(CSRC)/host.cu $(CSRC)/nve.cu $(CSRC)/track.cu $(CSRC)/debug.cu
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(CSRC)/cavb200_internal.cuh $(CSRC)/hotpath.cuh include/cavb200.h

all: lib oracle plugin

lib: $(LIBDIR)/libcavb200.so

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIBDIR)/libcavb200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) --no-print-directory -C oracle

plugin: lib
	@if [ -f plugin/Makefile ]; then $(MAKE) --no-print-directory -C plugin; fi

clean:
	rm -rf build $(LIBDIR)/libcavb200.so
	$(MAKE) -C oracle clean

.PHONY: all lib oracle plugin clean
