# Build of the three product libraries (+ the oracle, which is test infrastructure only).
#   ospray_b200/lib/libospcuda.so            CUDA kernels behind the C-ABI in include/ospcuda.h
#   ospray_b200/lib/libospray.so             OSPRay C API shim + device registry + module loader
#   ospray_b200/lib/libospray_module_cuda.so the `cuda` device (ospLoadModule("cuda"))
#   oracle/liboracle.so                      scalar CPU restatement (checker only)
NVCC      ?= nvcc
CXX       ?= g++
LIB       := ospray_b200/lib
CSRC      := ospray_b200/csrc
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
CXXFLAGS  := -std=c++17 -O2 -fPIC -Iinclude -Wall -Wno-unused-function
CUDA_LIBDIR ?= /usr/local/cuda/lib64

CU_SRCS   := $(CSRC)/cuda/runtime.cu $(CSRC)/cuda/render.cu $(CSRC)/cuda/fb_ops.cu $(CSRC)/cuda/bvh_build.cu $(CSRC)/cuda/peer_exchange.cu
CU_OBJS   := $(CU_SRCS:$(CSRC)/cuda/%.cu=build/cuda/%.o)
CU_HDRS   := $(wildcard $(CSRC)/cuda/*.cuh) $(wildcard $(CSRC)/cuda/*.h) $(CSRC)/cuda/tables.inc $(CSRC)/common/osp_math.h include/ospcuda.h
API_SRCS  := $(CSRC)/common/Managed.cpp $(CSRC)/api/Device.cpp $(CSRC)/api/API.cpp
DEV_SRCS  := $(wildcard $(CSRC)/device/*.cpp)
HOST_HDRS := $(wildcard $(CSRC)/common/*.h) $(wildcard $(CSRC)/api/*.h) $(wildcard $(CSRC)/device/*.h) $(wildcard include/ospray/*.h) include/ospcuda.h

all: $(LIB)/libospcuda.so $(LIB)/libospray.so $(LIB)/libospray_module_cuda.so oracle

build/cuda/%.o: $(CSRC)/cuda/%.cu $(CU_HDRS)
	@mkdir -p build/cuda
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB)/libospcuda.so: $(CU_OBJS)
	@mkdir -p $(LIB)
	$(NVCC) -shared -o $@ $(CU_OBJS) -lcudart

$(LIB)/libospray.so: $(API_SRCS) $(HOST_HDRS)
	@mkdir -p $(LIB)
	$(CXX) $(CXXFLAGS) -shared $(API_SRCS) -ldl -Wl,--exclude-libs,ALL -o $@

$(LIB)/libospray_module_cuda.so: $(DEV_SRCS) $(HOST_HDRS) $(LIB)/libospray.so $(LIB)/libospcuda.so
	$(CXX) $(CXXFLAGS) -shared $(DEV_SRCS) -L$(LIB) -lospray -lospcuda -lpthread -Wl,--exclude-libs,ALL -Wl,-rpath,'$$ORIGIN' -o $@

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)/*.so oracle/*.so oracle/_ref

.PHONY: all oracle clean
