# Builds the three native pieces in-tree:
#   feature_base_pointcloud_registration_b200/libfbpr_b200.so  -- the product: CUDA kernels + C ABI (sm_100a only)
#   synth/libsynth.so                                          -- synthetic workload generator (plain C)
#   oracle/liboracle.so                                        -- TEST INFRASTRUCTURE: CPU restatement of the reference
NVCC ?= /usr/local/cuda/bin/nvcc
PKG := feature_base_pointcloud_registration_b200
CSRC := $(PKG)/csrc
CU := $(CSRC)/capi.cu $(CSRC)/voxel.cu $(CSRC)/mapgrid.cu $(CSRC)/lm.cu $(CSRC)/projection.cu $(CSRC)/features.cu $(CSRC)/mapops.cu $(CSRC)/keyframes.cu $(CSRC)/selftest.cu
HDR := $(CSRC)/internal.cuh $(CSRC)/mapgrid.cuh $(CSRC)/smallmat.cuh include/fbpr_b200.h
# -fmad=false / -prec-div / -prec-sqrt: the kernels mirror the reference's f32 arithmetic op for op
NVFLAGS := $(EXTRA) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
           -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v

all: $(PKG)/libfbpr_b200.so synth/libsynth.so oracle/liboracle.so $(PKG)/host/libfeature_matching_b200.so

$(PKG)/libfbpr_b200.so: $(CU) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CU) -lcudart

$(PKG)/host/libfeature_matching_b200.so: $(PKG)/host/feature_matching.cpp $(PKG)/host/feature_matching.hpp include/fbpr_b200.h $(PKG)/libfbpr_b200.so
	/usr/bin/g++ -O2 -std=c++17 -fPIC -shared -Iinclude -o $@ $(PKG)/host/feature_matching.cpp -L$(PKG) -lfbpr_b200 -Wl,-rpath,'$$ORIGIN/..'

synth/libsynth.so: synth/synth.c
	/usr/bin/gcc -O2 -std=c11 -fPIC -shared -fvisibility=hidden -o $@ $< -lm

oracle/liboracle.so: oracle/oracle_capi.cpp oracle/ref_pipeline.hpp oracle/ref_cloud.hpp oracle/ref_smallmat.hpp
	$(MAKE) -C oracle liboracle.so

clean:
	rm -f $(PKG)/libfbpr_b200.so $(PKG)/host/*.so synth/libsynth.so oracle/liboracle.so
