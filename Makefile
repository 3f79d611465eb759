# Builds the C-ABI shared library of the hot path (sm_100a only); the oracle is pure PyTorch and has no native part.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# A/B experiments: `make BUILD=build_alt LIB=tcow_b200/libtcow_alt.so EXTRA=-DFOO=1`, then TCOW_B200_LIB=<that .so>.
EXTRA ?=
BUILD ?= build
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Xptxas -v $(EXTRA)
SRC := $(wildcard tcow_b200/csrc/*.cu)
OBJ := $(patsubst tcow_b200/csrc/%.cu,$(BUILD)/%.o,$(SRC))
LIB ?= tcow_b200/libtcow_b200.so

all: $(LIB)

HDR := $(wildcard tcow_b200/csrc/*.cuh) $(wildcard tcow_b200/csrc/*.h) include/tcow_b200.h

$(BUILD)/%.o: tcow_b200/csrc/%.cu $(HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; exit 1)

$(LIB): $(OBJ)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJ) -cudart static

clean:
	rm -rf $(BUILD) $(LIB)
.PHONY: all clean
