# Build of the B200-native scan engine.  Everything is built IN-TREE (the .so travels with gpurun).
#   make            -> trre_b200/lib/libtrre_b200.so, bin/trre, bin/trre_dft, oracle (checker)
NVCC    ?= /usr/local/cuda/bin/nvcc
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unused-parameter
CSRC    := trre_b200/csrc
LIB     := trre_b200/lib/libtrre_b200.so
SRCS    := $(CSRC)/engine.cu $(CSRC)/pattern.cpp $(CSRC)/automaton.cpp $(CSRC)/host_extras.cpp
HDRS    := $(wildcard $(CSRC)/*.hpp $(CSRC)/*.cuh) include/trre_b200.h

all: $(LIB) cli oracle

$(LIB): $(SRCS) $(HDRS)
	@mkdir -p trre_b200/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRCS)

cli: bin/trre bin/trre_dft

bin/trre: $(CSRC)/cli_main.cpp $(LIB)
	@mkdir -p bin
	g++ -O2 -std=c++17 -pthread -Wall -Wextra -DTRRE_CLI_FLAVOR=0 $< -o $@ -Ltrre_b200/lib -ltrre_b200 -Wl,-rpath,'$$ORIGIN/../trre_b200/lib'

bin/trre_dft: $(CSRC)/cli_main.cpp $(LIB)
	@mkdir -p bin
	g++ -O2 -std=c++17 -pthread -Wall -Wextra -DTRRE_CLI_FLAVOR=1 $< -o $@ -Ltrre_b200/lib -ltrre_b200 -Wl,-rpath,'$$ORIGIN/../trre_b200/lib'

oracle:
	$(MAKE) -s -C oracle

ptxas-info:
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -Xptxas -v -c $(CSRC)/engine.cu -o /dev/null

clean:
	rm -rf trre_b200/lib bin
	$(MAKE) -C oracle clean

.PHONY: all cli oracle clean ptxas-info
