# Builds the sm_100a C-ABI library in-tree (the .so is git-ignored but travels to the GPU box).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard ai_toolkit_b200/csrc/*.cu)
OBJ := $(patsubst ai_toolkit_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := ai_toolkit_b200/lib/libb200lora.so

LIB_PDL := ai_toolkit_b200/lib/libb200lora_pdl.so
OBJ_PDL := $(patsubst ai_toolkit_b200/csrc/%.cu,build_pdl/%.o,$(SRC))

all: $(LIB) $(LIB_PDL)

# opt-in variant: every launch carries the programmatic-stream-serialization attribute and every kernel waits with
# griddepcontrol (common.cuh).  Not the default: select with B200_LIB=ai_toolkit_b200/lib/libb200lora_pdl.so
pdl: $(LIB_PDL)

build_pdl/%.o: ai_toolkit_b200/csrc/%.cu $(wildcard ai_toolkit_b200/csrc/*.cuh) $(wildcard ai_toolkit_b200/csrc/*.h) include/b200_lora.h
	@mkdir -p build_pdl
	$(NVCC) $(NVCCFLAGS) -DB200_PDL=1 -c $< -o $@ 2> build_pdl/$*.ptxas.log || (cat build_pdl/$*.ptxas.log; exit 1)

$(LIB_PDL): $(OBJ_PDL)
	@mkdir -p ai_toolkit_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ_PDL) -lcudart_static -ldl -lpthread -lrt

build/%.o: ai_toolkit_b200/csrc/%.cu $(wildcard ai_toolkit_b200/csrc/*.cuh) $(wildcard ai_toolkit_b200/csrc/*.h) include/b200_lora.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p ai_toolkit_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -ldl -lpthread -lrt

# tuning variants for same-box A/B runs:  make variant NAME=poly0 DEFS="-DB200_ATTN_POLY_R2=0"
#   -> ai_toolkit_b200/lib/libb200lora_poly0.so (select with B200_LIB)
variant:
	@test -n "$(NAME)" || (echo "usage: make variant NAME=<tag> DEFS=\"-D...\""; exit 1)
	@mkdir -p build_$(NAME) ai_toolkit_b200/lib
	for f in $(SRC); do $(NVCC) $(NVCCFLAGS) $(DEFS) -c $$f -o build_$(NAME)/$$(basename $$f .cu).o 2> build_$(NAME)/$$(basename $$f .cu).ptxas.log || exit 1; done
	$(NVCC) $(ARCH) -shared -o ai_toolkit_b200/lib/libb200lora_$(NAME).so build_$(NAME)/*.o -lcudart_static -ldl -lpthread -lrt

clean:
	rm -rf build build_pdl $(LIB) $(LIB_PDL)
.PHONY: all pdl variant clean
