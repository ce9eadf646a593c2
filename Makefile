# Builds libdiffsound_b200.so (sm_100a only) in-tree; the .so travels to the GPU box with the gpurun snapshot.
NVCC      ?= /usr/local/cuda/bin/nvcc
PKG       := text-to-sound-synthesis_b200
CSRC      := $(PKG)/csrc
OUT       := $(PKG)/libdiffsound_b200.so
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -I$(CSRC) \
             --expt-relaxed-constexpr -Xptxas -v

all: $(OUT)

build/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/diffsound_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(OUT): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart

clean:
	rm -rf build $(OUT)
.PHONY: all clean
