# libfesom2-accelerate.so for NVIDIA B200 (sm_100a).  Replaces the reference's Makefile
# (/root/reference/Makefile: one -rdc object per kernel for sm_60); here the kernels are header
# templates instantiated by the driver, so there are five translation units and no device linking.
NVCC     ?= nvcc
ARCH     ?= -gencode arch=compute_100a,code=sm_100a
# -fmad=false: no FMA contraction, so kernels round exactly like the CPU reference (g++ without -march)
NVFLAGS  = -std=c++17 -O3 -lineinfo $(ARCH) -fmad=false -Xcompiler -fPIC,-Wall,-Wno-unused-function
PKG      = fesom2-accelerate_b200
SRC      = $(PKG)/csrc
OUT      = $(PKG)/lib
OBJS     = $(OUT)/fct_driver.o $(OUT)/fct_fields.o $(OUT)/fct_halo.o $(OUT)/fct_plan.o $(OUT)/stress2rhs.o
HDRS     = $(SRC)/fct_kernels.cuh $(SRC)/fct_tile_kernels.cuh $(SRC)/fct_warp_kernels.cuh $(SRC)/fct_plan.h $(SRC)/fct_internal.h include/fesom2-accelerate.h

all: $(OUT)/libfesom2-accelerate.so oracle

$(OUT)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OUT)
	$(NVCC) $(NVFLAGS) $(PTXAS) -c $< -o $@

$(OUT)/libfesom2-accelerate.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(OUT)/*.o $(OUT)/*.so
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
