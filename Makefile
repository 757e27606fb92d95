# Native build of the C-ABI library without Python (same commands as __graft_entry__.build()).
NVCC      ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
CSRC      := biogo_b200/csrc
OBJDIR    := biogo_b200/lib/obj
LIB       := biogo_b200/lib/libbiogoalign.so
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h $(CSRC)/*.inl) include/biogo_align.h
FAMILIES  := 0_0 1_0 2_0 0_1 1_1 2_1 0_2 1_2 2_2
OBJS      := $(OBJDIR)/ba_api.o $(foreach f,$(FAMILIES),$(OBJDIR)/ba_fill16_$(f).o)

all: $(LIB) biogo_b200/lib/ba_microbench

$(OBJDIR):
	mkdir -p $(OBJDIR)

$(OBJDIR)/ba_api.o: $(CSRC)/ba_api.cu $(HDRS) | $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

# one translation unit per (kind, linear) family of the packed fill kernel: make -j7 builds them in parallel
$(OBJDIR)/ba_fill16_%.o: $(CSRC)/ba_fill16_inst.cu $(HDRS) | $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -DBA_INST_KIND=$(word 1,$(subst _, ,$*)) -DBA_INST_LIN=$(word 2,$(subst _, ,$*)) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(OBJS)

biogo_b200/lib/ba_microbench: $(CSRC)/ba_microbench.cu | $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -o $@ $<

clean:
	rm -rf biogo_b200/lib tests/emu/_build oracle/_build

.PHONY: all clean
