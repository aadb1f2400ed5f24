# Builds the C-ABI shared library (sm_100a only) in-tree.  `make` is what __graft_entry__.build() runs.
NVCC      ?= nvcc
PKG       := uda_aerial_semantic_segmentation_research_b200
CSRC      := $(PKG)/csrc
BUILD     := build
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB       := $(PKG)/libuda_b200.so

all: $(LIB) oracle

$(BUILD)/%.o: $(CSRC)/%.cu $(CSRC)/common.cuh include/uda_b200.h $(wildcard $(CSRC)/*.cuh)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -lcudart

oracle: oracle/_build/libuda_oracle.so

oracle/_build/libuda_oracle.so: oracle/confmat_oracle.c
	@mkdir -p oracle/_build
	gcc -O2 -shared -fPIC -o $@ $<

clean:
	rm -rf $(BUILD) $(LIB) oracle/_build

.PHONY: all clean oracle
