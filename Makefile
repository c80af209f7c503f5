# Builds the C-ABI shared library (sm_100a only) and the oracle helpers.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
SRC_DIR   := tricolo_b200/csrc
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,build/%.o,$(SRCS))
LIB       := tricolo_b200/lib/libtricolo_b200.so

all: $(LIB)

build/%.o: $(SRC_DIR)/%.cu $(wildcard $(SRC_DIR)/*.h) $(wildcard $(SRC_DIR)/*.cuh) include/tricolo_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIB): $(OBJS)
	@mkdir -p tricolo_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -cudart shared

# wait-time accounting build (profiles/pc_trace.py, profiles/fwd_trace.py)
TRACE_LEVEL ?= 2
EXP ?= 0
TRACE_LIB ?= tricolo_b200/lib/libtricolo_b200_trace.so
TRACE_DEFS ?=
trace:
	@mkdir -p build_trace
	for f in $(SRCS); do $(NVCC) $(NVCCFLAGS) -DTCL_PAIR_TRACE=$(TRACE_LEVEL) -DTCL_PAIR_EXP=$(EXP) $(TRACE_DEFS) -c $$f -o build_trace/$$(basename $$f .cu).o 2> /dev/null || exit 1; done
	$(NVCC) $(ARCH) -shared -o $(TRACE_LIB) build_trace/*.o -cudart shared

clean:
	rm -rf build build_trace $(LIB) tricolo_b200/lib/libtricolo_b200_trace.so

.PHONY: all clean trace
