# Builds the product library (CUDA kernels + C ABI + host pipeline) and the test-only CPU oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
GENCODE := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(GENCODE) -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Wno-sign-compare
CSRC := glnexus_b200/csrc
HOST_SRCS := $(CSRC)/glx_host.cc $(CSRC)/glx_host_c.cc $(CSRC)/glx_presets.cc $(CSRC)/glx_synth.cc $(CSRC)/glx_service.cc $(CSRC)/glx_bcf_host.cc $(CSRC)/glx_stream.cc $(CSRC)/glx_wire_host.cc
HOST_OBJS := $(HOST_SRCS:.cc=.o)

all: glnexus_b200/libglx.so oracle/libglx_oracle.so

$(CSRC)/%.o: $(CSRC)/%.cc $(CSRC)/glx_host.hpp include/glx.h include/glx_host.h
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(CSRC)/glx_device.o: $(CSRC)/glx_device.cu $(wildcard $(CSRC)/*.cuh) include/glx.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(CSRC)/glx_device.ptxas.log || (cat $(CSRC)/glx_device.ptxas.log; false)

$(CSRC)/glx_bcf.o: $(CSRC)/glx_bcf.cu $(wildcard $(CSRC)/*.cuh) include/glx.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(CSRC)/glx_bcf.ptxas.log || (cat $(CSRC)/glx_bcf.ptxas.log; false)

glnexus_b200/libglx.so: $(HOST_OBJS) $(CSRC)/glx_device.o $(CSRC)/glx_bcf.o
	$(NVCC) -shared $(GENCODE) -o $@ $^ -cudart static -lpthread

oracle/libglx_oracle.so: oracle/glx_oracle.cc oracle/glx_bcf_oracle.cc include/glx.h
	$(CXX) -O3 -march=x86-64-v3 -std=c++17 -fPIC -shared -o $@ oracle/glx_oracle.cc oracle/glx_bcf_oracle.cc -lpthread

clean:
	rm -f $(CSRC)/*.o glnexus_b200/libglx.so oracle/libglx_oracle.so $(CSRC)/*.ptxas.log

# A/B builds of the same sources with extra defines (ablations, experiments): make alt NAME=nopf DEFS=-DGLX_ABLATE_ROW_PREFETCH
# -> build_alt/libglx_$(NAME).so, loaded with GLX_LIB_PATH=build_alt/libglx_$(NAME).so
alt: $(HOST_OBJS)
	mkdir -p build_alt
	$(NVCC) $(NVCCFLAGS) $(DEFS) -c $(CSRC)/glx_device.cu -o build_alt/glx_device_$(NAME).o 2> build_alt/glx_device_$(NAME).ptxas.log
	$(NVCC) $(NVCCFLAGS) $(DEFS) -c $(CSRC)/glx_bcf.cu -o build_alt/glx_bcf_$(NAME).o 2> build_alt/glx_bcf_$(NAME).ptxas.log
	$(NVCC) -shared $(GENCODE) -o build_alt/libglx_$(NAME).so $(HOST_OBJS) build_alt/glx_device_$(NAME).o build_alt/glx_bcf_$(NAME).o -cudart static -lpthread
	rm -f build_alt/*.o
