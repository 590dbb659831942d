# Build of the B200 DSV2 pixel path.
#
#   make            -> libdsv2cuda.so (CUDA, sm_100a; the product) + dsv2cu CLI
#   make emu        -> tests/_emu/libdsv2cuda_emu.so (kernel sources compiled for
#                      the host; TEST-ONLY, lets the CPU test suite exercise the
#                      kernel arithmetic in a container without a GPU)
#   make oracle     -> oracle/liboracle.so (CPU restatement, test infrastructure)
#   make ref        -> oracle/_ref/ (the unmodified reference, built from
#                      /root/reference when present; test infrastructure)
PKG     := digital-subband-video-2_b200
CSRC    := $(PKG)/csrc
HOST    := $(PKG)/host
NVCC    ?= nvcc
CC      ?= gcc
CXX     ?= g++
ARCH    := -gencode arch=compute_100a,code=sm_100a
EXTRA   ?=
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default $(EXTRA)
CFLAGS  := -O2 -fPIC -Wall -Wextra -Wno-unused-parameter -Iinclude
HOSTSRC := $(HOST)/dsv_core.c $(HOST)/dsv_bits.c $(HOST)/dsv_hzcc.c $(HOST)/dsv_mvutil.c \
           $(HOST)/dsv_dec.c $(wildcard $(HOST)/dsv_enc.c) $(wildcard $(HOST)/dsv_ops.c) \
           $(wildcard $(HOST)/dsv_pipe.c)
HOSTOBJ := $(HOSTSRC:.c=.o)
EMUOBJ  := $(HOSTSRC:.c=.emu.o)
KHDRS   := $(wildcard $(CSRC)/*.cuh) $(CSRC)/dsvcu_rt.h include/dsv_cuda.h
REFSRC  := /root/reference/src

CLI := $(if $(wildcard $(HOST)/dsv_cli.c),$(PKG)/dsv2cu,)
all: $(PKG)/libdsv2cuda.so $(CLI)

$(CSRC)/dsvcu_api.o: $(CSRC)/dsvcu_api.cu $(KHDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(HOST)/%.o: $(HOST)/%.c $(HOST)/dsv_host.h include/dsv.h include/dsv_cuda.h
	$(CC) $(CFLAGS) -c $< -o $@

$(PKG)/libdsv2cuda.so: $(CSRC)/dsvcu_api.o $(HOSTOBJ)
	$(NVCC) $(ARCH) -shared -o $@ $^ -Xlinker -Bsymbolic -lpthread

$(PKG)/dsv2cu: $(HOST)/dsv_cli.c $(PKG)/libdsv2cuda.so
	$(CC) $(CFLAGS) -o $@ $(HOST)/dsv_cli.c -L$(PKG) -ldsv2cuda -Wl,-rpath,'$$ORIGIN' -lpthread

# ---- test-only host emulation of the kernel sources
emu: tests/_emu/libdsv2cuda_emu.so
$(HOST)/%.emu.o: $(HOST)/%.c $(HOST)/dsv_host.h include/dsv.h include/dsv_cuda.h
	$(CC) $(CFLAGS) -c $< -o $@
tests/_emu/dsvcu_api_emu.o: $(CSRC)/dsvcu_api.cu $(KHDRS)
	mkdir -p tests/_emu
	$(CXX) -x c++ -O2 -fPIC -DDSVCU_EMU -c $< -o $@
tests/_emu/libdsv2cuda_emu.so: tests/_emu/dsvcu_api_emu.o $(EMUOBJ)
	$(CXX) -shared -o $@ $^ -Wl,-Bsymbolic -lpthread

# ---- the unmodified reference (only where /root/reference exists)
ref:
	@if [ -d $(REFSRC) ]; then $(MAKE) -C oracle ref; else echo "no $(REFSRC): using prebuilt oracle/_ref"; fi
oracle:
	$(MAKE) -C oracle liboracle.so

clean:
	rm -f $(CSRC)/*.o $(HOST)/*.o $(PKG)/libdsv2cuda.so $(PKG)/dsv2cu tests/_emu/*.o tests/_emu/*.so

.PHONY: all emu ref oracle clean
