# Builds the sm_100a C-ABI library in-tree (the .so is git-ignored but travels to the GPU box).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr
SRC := $(wildcard ai_toolkit_b200/csrc/*.cu)
OBJ := $(patsubst ai_toolkit_b200/csrc/%.cu,build/%.o,$(SRC))
LIB := ai_toolkit_b200/lib/libb200lora.so

all: $(LIB)

build/%.o: ai_toolkit_b200/csrc/%.cu $(wildcard ai_toolkit_b200/csrc/*.cuh) $(wildcard ai_toolkit_b200/csrc/*.h) include/b200_lora.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	@mkdir -p ai_toolkit_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart_static -ldl -lpthread -lrt

clean:
	rm -rf build $(LIB)
.PHONY: all clean
