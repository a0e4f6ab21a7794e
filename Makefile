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
PLUGIN_SRCS := $(CSRC)/plugin.cu $(CSRC)/pathtrace.cu $(CSRC)/mesh_trace.cu $(CSRC)/mesh_wavefront.cu $(CSRC)/lbvh_build.cu $(CSRC)/post.cu
PLUGIN_HDRS := $(CSRC)/mesh_traverse.cuh $(CSRC)/quantized_bvh.hpp $(CSRC)/kernels.hpp $(CSRC)/scene_params.hpp $(CSRC)/scan_decide.hpp $(CSRC)/slab_test.hpp $(CSRC)/wide_bvh.hpp $(CSRC)/sincos_table.hpp $(CSRC)/sincos_table_data.inc $(CSRC)/sincos_table_constants.inc $(CSRC)/device_common.cuh $(CSRC)/mesh_scene.hpp include/quasi_plugin_abi.h \
               include/quasi_cuda_backend.h
PLUGIN_SO := quasi_b200/lib/libquasi_cuda.so

all: plugin host oracle

plugin: $(PLUGIN_SO)

$(PLUGIN_SO): $(PLUGIN_SRCS) $(PLUGIN_HDRS)
	mkdir -p $(dir $@)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(PLUGIN_SRCS)

# ---- C++ headless host (g++ -std=c++23, the reference's .bazelrc flags) ----
HOST := quasi_b200/host
CXX ?= g++
HOST_CXXFLAGS := -std=c++23 -O2 -Wall -Wextra -Wpedantic -fcoroutines -pthread -fvisibility=hidden -fPIC \
                 -ffp-contract=off -I$(HOST)/include -Iinclude -I/usr/local/cuda/include
HOST_HDRS := $(shell find $(HOST)/include -name '*.hpp') include/quasi_plugin_abi.h include/quasi_cuda_backend.h
HOST_BIN := quasi_b200/bin/quasi_headless

host: $(HOST_BIN)

$(HOST_BIN): $(HOST)/src/main.cpp $(HOST)/src/cuda_context.cpp $(HOST)/src/exr_writer.cpp $(HOST_HDRS)
	mkdir -p $(dir $@)
	$(CXX) $(HOST_CXXFLAGS) -o $@ $(HOST)/src/main.cpp $(HOST)/src/cuda_context.cpp $(HOST)/src/exr_writer.cpp \
		-L/usr/local/cuda/lib64 -lcudart -lz -ldl

# C++ tests: this repository's host tests (with a fake plugin fixture) and, where
# /root/reference exists, the reference's own three test files compiled unmodified
# from where they lie against this repository's headers.
TESTDIR := quasi_b200/bin/tests
# ASan + UBSan where the toolchain ships the runtime (some g++ installs do not).
SANITIZE := $(shell echo 'int main(){}' | $(CXX) -fsanitize=address,undefined -x c++ - -o /dev/null 2>/dev/null \
              && echo -fsanitize=address,undefined)
TEST_CXXFLAGS := $(HOST_CXXFLAGS) -O1 -g $(SANITIZE) -I$(HOST)/test
FAKE_SRC := $(HOST)/test/fake_plugin.cpp

host-tests: $(TESTDIR)/host_test $(TESTDIR)/exr_write_tool $(TESTDIR)/mesh_dump_tool

$(TESTDIR)/mesh_dump_tool: $(HOST)/test/mesh_dump_tool.cpp quasi_b200/csrc/mesh_scene.hpp
	mkdir -p $(TESTDIR)
	$(CXX) -std=c++20 -O2 -ffp-contract=off -Wall -Wextra -pthread -Iquasi_b200/csrc -o $@ $(HOST)/test/mesh_dump_tool.cpp
	if [ -d $(REFERENCE)/test ]; then $(MAKE) reference-tests; fi

$(TESTDIR)/host_test: $(HOST)/test/host_test.cpp $(HOST)/test/catch2_main.cpp $(HOST)/src/exr_writer.cpp \
		$(FAKE_SRC) $(HOST_HDRS) $(HOST)/test/catch2/catch_test_macros.hpp
	mkdir -p $(TESTDIR)
	$(CXX) -std=c++20 -O1 -fPIC -shared -Iinclude -DFAKE_PLUGIN_GENERATION=1 $(FAKE_SRC) -o $(TESTDIR)/libfake_v1.so
	$(CXX) -std=c++20 -O1 -fPIC -shared -Iinclude -DFAKE_PLUGIN_GENERATION=2 $(FAKE_SRC) -o $(TESTDIR)/libfake_v2.so
	$(CXX) -std=c++20 -O1 -fPIC -shared -Iinclude -DFAKE_PLUGIN_ABI=99 $(FAKE_SRC) -o $(TESTDIR)/libfake_bad_abi.so
	$(CXX) -std=c++20 -O1 -fPIC -shared -Iinclude -DFAKE_PLUGIN_NO_READBACK $(FAKE_SRC) -o $(TESTDIR)/libfake_minimal.so
	$(CXX) $(TEST_CXXFLAGS) -o $@ $(HOST)/test/host_test.cpp $(HOST)/test/catch2_main.cpp \
		$(HOST)/src/exr_writer.cpp -lz -ldl

$(TESTDIR)/exr_write_tool: $(HOST)/test/exr_write_tool.cpp $(HOST)/src/exr_writer.cpp $(HOST_HDRS)
	mkdir -p $(TESTDIR)
	$(CXX) $(TEST_CXXFLAGS) -o $@ $(HOST)/test/exr_write_tool.cpp $(HOST)/src/exr_writer.cpp -lz

# ref_loader_host: tests/ref_loader_host.cpp against the REFERENCE's own loader / manager headers
# (-I$(REFERENCE)/src, none of this repository's host headers): the drop-in proof.
reference-tests:
	mkdir -p $(TESTDIR)
	$(CXX) -std=c++23 -O1 -Wall -Wextra -Wpedantic -fcoroutines -pthread -I$(REFERENCE)/src -Iinclude \
		-I/usr/local/cuda/include -o $(TESTDIR)/ref_loader_host tests/ref_loader_host.cpp \
		-L/usr/local/cuda/lib64 -lcudart -ldl
	for t in async_test plugin_test exr_writer_test; do \
		$(CXX) $(TEST_CXXFLAGS) -o $(TESTDIR)/ref_$$t $(REFERENCE)/test/$$t.cpp $(HOST)/test/catch2_main.cpp \
			$(HOST)/src/exr_writer.cpp -lz -ldl || exit 1; \
	done

oracle:
	$(MAKE) -C oracle all
	if [ -d $(REFERENCE) ]; then $(MAKE) -C oracle ref REFERENCE=$(REFERENCE); fi

ptxas-info:
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -shared -o /tmp/libquasi_cuda_v.so $(PLUGIN_SRCS)

clean:
	rm -rf quasi_b200/lib quasi_b200/bin
	$(MAKE) -C oracle clean

.PHONY: all plugin host host-tests reference-tests oracle clean ptxas-info
