# Builds the C-ABI CUDA library (sm_100a only) and the C oracle.  `python -c "import __graft_entry__ as g; g.build()"` runs this.
NVCC ?= nvcc
CSRC := arm_pose_estimation_b200/csrc
FKFLAGS ?=
LIB  := arm_pose_estimation_b200/lib/libape_b200.so
SRCS := $(wildcard $(CSRC)/*.cu)
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/ape_b200.h
NVFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude

all: $(LIB)

$(LIB): $(SRCS) $(HDRS)
	mkdir -p $(dir $(LIB))
	$(NVCC) -shared $(NVFLAGS) -o $@ $(SRCS)

ptxas-info: $(SRCS) $(HDRS)
	$(NVCC) -shared $(NVFLAGS) -Xptxas -v -o /dev/null $(SRCS)

clean:
	rm -f $(LIB)

.PHONY: all clean ptxas-info
