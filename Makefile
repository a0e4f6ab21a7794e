# Build of the CUDA backend and its checker. `bazel` is not available in this
# environment; the BUILD.bazel files next to the sources describe
# the same targets for a tree that has rules_cuda.
#
#   make            libquasi_cuda.so + the CPU oracle
#   make plugin     quasi_b200/lib/libquasi_cuda.so only
#   make oracle     oracle/libquasi_oracle.so (+ oracle/_ref where /root/reference exists)

NVCC ?= /usr/local/cuda/bin/nvcc
REFERENCE ?= /root/reference

ARCH := -gencode arch=compute_100a,code=sm_100a
# Numerics contract: no FMA contraction, IEEE div/sqrt, denormals kept.
NVCC_NUMERICS := -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
HOST_FLAGS := -fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall,-Wextra
NVCCFLAGS := $(ARCH) -O3 -std=c++20 -lineinfo $(NVCC_NUMERICS) --cudart shared \
             -Xcompiler $(HOST_FLAGS) -Iinclude -Iquasi_b200/csrc

CSRC := quasi_b200/csrc
PLUGIN_SRCS := $(CSRC)/plugin.cu $(CSRC)/pathtrace.cu $(CSRC)/post.cu
PLUGIN_HDRS := $(CSRC)/kernels.hpp $(CSRC)/scene_params.hpp include/quasi_plugin_abi.h \
               include/quasi_cuda_backend.h
PLUGIN_SO := quasi_b200/lib/libquasi_cuda.so

all: plugin oracle

plugin: $(PLUGIN_SO)

$(PLUGIN_SO): $(PLUGIN_SRCS) $(PLUGIN_HDRS)
	mkdir -p $(dir $@)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(PLUGIN_SRCS)

oracle:
	$(MAKE) -C oracle all
	if [ -d $(REFERENCE) ]; then $(MAKE) -C oracle ref REFERENCE=$(REFERENCE); fi

ptxas-info:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared -o /tmp/libquasi_cuda_v.so $(PLUGIN_SRCS)

clean:
	rm -rf quasi_b200/lib
	$(MAKE) -C oracle clean

.PHONY: all plugin oracle clean ptxas-info
