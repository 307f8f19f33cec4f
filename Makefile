# Builds libparseable_b200.so (CUDA, sm_100a only) in-tree, the C oracle and the
# CPU test harness for the pure decode functions.  No PyTorch, no Triton.
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
PY_NCCL   := $(shell python -c "import nvidia.nccl,os;print(os.path.dirname(nvidia.nccl.__file__))" 2>/dev/null)
NCCL_INC  := $(if $(PY_NCCL),-I$(PY_NCCL)/include,)
# link the torch-bundled libnccl.so.2 when present (same soname as the system one, so a
# process that also imports torch ends up with a single NCCL)
NCCL_LIB  := $(if $(PY_NCCL),-L$(PY_NCCL)/lib -l:libnccl.so.2 -Xlinker -rpath -Xlinker $(PY_NCCL)/lib,-lnccl)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $(NCCL_INC) -Iinclude $(EXTRA)
CSRC      := parseable_b200/csrc
OBJDIR    ?= build
LIB       ?= parseable_b200/libparseable_b200.so
EXTRA     ?=

CU_SRCS   := $(CSRC)/table.cu $(CSRC)/query.cu
CPP_SRCS  := $(CSRC)/parquet_meta.cpp $(CSRC)/arrow_export.cpp $(CSRC)/capi.cpp $(CSRC)/comm.cpp $(CSRC)/planning.cpp
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS)) $(patsubst $(CSRC)/%.cpp,$(OBJDIR)/%.o,$(CPP_SRCS))
HDRS      := $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh include/*.h)

all: $(LIB) oracle tools

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

$(OBJDIR)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart $(NCCL_LIB)

oracle: oracle/liboracle.so
oracle/liboracle.so: oracle/oracle.c
	$(CC) -O2 -std=c11 -fPIC -shared -Wall -o $@ $< -lm

tools: tools/libdecode_core_host.so tools/libzstd_host.so tools/libjson_host.so
tools/libjson_host.so: tools/json_host.cpp $(CSRC)/json_egress.cuh $(CSRC)/ryu_f64.cuh $(CSRC)/ryu_tables.inc
	$(CXX) -O2 -std=c++17 -fPIC -shared -Wall -Wno-maybe-uninitialized -I$(CSRC) -o $@ $<
tools/libzstd_host.so: tools/zstd_host.cpp $(CSRC)/zstd_decode.cuh $(CSRC)/inflate_decode.cuh
	$(CXX) -O2 -std=c++17 -fPIC -shared -Wall -I$(CSRC) -o $@ $<
tools/libdecode_core_host.so: tools/decode_core_host.cpp $(CSRC)/decode_core.cuh $(CSRC)/device_structs.hpp
	$(CXX) -O2 -std=c++17 -fPIC -shared -Wall -I$(CSRC) -o $@ $<

clean:
	rm -rf $(OBJDIR) $(LIB) oracle/liboracle.so tools/libdecode_core_host.so tools/libzstd_host.so tools/libjson_host.so tools/libjson_host.so

.PHONY: all oracle tools clean
