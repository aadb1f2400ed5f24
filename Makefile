# Builds the C-ABI shared library (sm_100a only) in-tree.  `make` is what __graft_entry__.build() runs.
NVCC      ?= nvcc
PKG       := uda_aerial_semantic_segmentation_research_b200
CSRC      := $(PKG)/csrc
BUILD     := build
LIBNAME   := libuda_b200.so
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function --expt-relaxed-constexpr
# make EXPERIMENTS=1 compiles the work-skipping timing switches (UDA_B200_TC_DEBUG / UDA_B200_BN_DEBUG) into an
# experiment build; the default product library does not contain them
ifeq ($(EXPERIMENTS),1)
NVCCFLAGS += -DUDA_B200_EXPERIMENTS
BUILD     := build_exp
LIBNAME   := libuda_b200_exp.so
endif
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SRCS))
LIB       := $(PKG)/$(LIBNAME)

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
