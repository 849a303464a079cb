# hast-b200 build.  Everything is built in-tree so that the artefacts travel
# with the gpurun snapshot:
#   hast_b200/lib/libhast_b200.so   the C-ABI engine (CUDA, sm_100a only)
#   hast_b200/lib/libhast_tools.so  synthetic-data helper (FASTQ text emitter)
#   bin/classify, bin/mergeResult   drop-in replacements for the reference stage-01 binaries
#   oracle/liboracle.so, oracle/_ref/*   test infrastructure (see oracle/Makefile)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  ?= g++
CC   ?= gcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function
CUDA_HOME ?= /usr/local/cuda

LIB := hast_b200/lib/libhast_b200.so
TOOLS := hast_b200/lib/libhast_tools.so
CSRC := hast_b200/csrc
HOST := hast_b200/host

.PHONY: all lib tools host oracle clean sass
all: lib tools host oracle

lib: $(LIB)
$(LIB): $(CSRC)/hast_b200.cu $(CSRC)/kcount.cuh $(CSRC)/kernels.cuh $(CSRC)/fused.cuh $(CSRC)/table.cuh $(CSRC)/kmer.cuh $(CSRC)/host_pack.cpp $(CSRC)/host_pack.h include/hast_b200.h
	@mkdir -p hast_b200/lib
	$(CXX) -O3 -std=c++17 -fPIC -Wall -pthread -c $(CSRC)/host_pack.cpp -o hast_b200/lib/host_pack.o
	$(NVCC) $(NVFLAGS) -Xptxas -v -shared $< hast_b200/lib/host_pack.o -o $@ -ldl 2> hast_b200/lib/ptxas.log || (cat hast_b200/lib/ptxas.log; exit 1)
	@rm -f hast_b200/lib/host_pack.o
	@grep -E "registers|spill" hast_b200/lib/ptxas.log | sort | uniq -c | sort -rn | head -20 || true

# synthetic-data helpers (FASTQ text emitter, counter-based read-pair generator: host and device builds)
SYNTH_CUDA := hast_b200/lib/libhast_synth_cuda.so
tools: $(TOOLS) $(SYNTH_CUDA)
$(TOOLS): hast_b200/tools/fastq_fmt.c hast_b200/tools/synth_gen_cpu.cpp hast_b200/tools/synth_gen.h
	@mkdir -p hast_b200/lib
	$(CC) -O2 -std=c11 -fPIC -c hast_b200/tools/fastq_fmt.c -o hast_b200/lib/fastq_fmt.o
	$(CXX) -O2 -std=c++17 -fPIC -pthread -shared hast_b200/tools/synth_gen_cpu.cpp hast_b200/lib/fastq_fmt.o -lz -o $@
	@rm -f hast_b200/lib/fastq_fmt.o
$(SYNTH_CUDA): hast_b200/tools/synth_gen.cu hast_b200/tools/synth_gen.h
	@mkdir -p hast_b200/lib
	$(NVCC) $(NVFLAGS) -shared $< -o $@

host: bin/classify bin/mergeResult bin/quartering_fastq bin/classify_seq bin/build_unshared_kmers bin/hast_gunzip bin/h2d_probe
HOST_SRCS := $(wildcard $(HOST)/*.cpp)
HOST_HDRS := $(wildcard $(HOST)/*.h)
MAINS := $(HOST)/merge_result_main.cpp $(HOST)/quartering_main.cpp $(HOST)/classify_main.cpp $(HOST)/classify_seq_main.cpp $(HOST)/build_unshared_main.cpp $(HOST)/gunzip_main.cpp
bin/classify: $(filter-out $(MAINS),$(HOST_SRCS)) $(HOST)/classify_main.cpp $(HOST_HDRS) $(LIB) include/hast_b200.h
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -pthread -Iinclude $(filter-out $(MAINS),$(HOST_SRCS)) $(HOST)/classify_main.cpp \
	    -Lhast_b200/lib -lhast_b200 -lz -Wl,-rpath,'$$ORIGIN/../hast_b200/lib' -o $@
# stage 03's per-sequence classifier on the same engine
bin/classify_seq: $(HOST)/classify_seq_main.cpp $(LIB) include/hast_b200.h
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -Iinclude $(HOST)/classify_seq_main.cpp -Lhast_b200/lib -lhast_b200 -lz \
	    -Wl,-rpath,'$$ORIGIN/../hast_b200/lib' -o $@
# stage 00's parent-unique k-mer lists from parental reads (count table on the GPU)
bin/build_unshared_kmers: $(HOST)/build_unshared_main.cpp $(HOST)/inflate.cpp $(HOST)/inflate.h $(HOST)/crc32_clmul.h $(LIB) include/hast_b200.h
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -pthread -Iinclude $(HOST)/build_unshared_main.cpp $(HOST)/inflate.cpp $(HOST)/inflate_par.cpp -Lhast_b200/lib -lhast_b200 -lz \
	    -Wl,-rpath,'$$ORIGIN/../hast_b200/lib' -o $@
# the read partitioner alone needs no GPU and no CUDA library
bin/quartering_fastq: $(HOST)/quartering_main.cpp $(HOST)/partition.cpp $(HOST)/fastq_source.cpp $(HOST_HDRS)
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -pthread -Iinclude $(HOST)/quartering_main.cpp $(HOST)/partition.cpp $(HOST)/fastq_source.cpp $(HOST)/plain_slicer.cpp $(HOST)/parser.cpp $(HOST)/barcode_index.cpp $(HOST)/inflate.cpp $(HOST)/inflate_par.cpp -lz -o $@
# the gzip decoder of the readers as a stand-alone tool (tests, timing)
bin/hast_gunzip: $(HOST)/gunzip_main.cpp $(HOST)/inflate.cpp $(HOST)/inflate_par.cpp $(HOST)/inflate.h $(HOST)/crc32_clmul.h $(HOST)/inflate_par.h
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall -pthread $(HOST)/gunzip_main.cpp $(HOST)/inflate.cpp $(HOST)/inflate_par.cpp -lz -o $@
# measurement tool: the box's host->device ceiling with 1/2/4/8 GPUs copying at once
bin/h2d_probe: hast_b200/tools/h2d_probe.cu
	@mkdir -p bin
	$(NVCC) $(ARCH) -O2 -std=c++17 -Xcompiler -pthread $< -o $@
bin/mergeResult: $(HOST)/merge_result_main.cpp
	@mkdir -p bin
	$(CXX) -O2 -g -std=c++17 -Wall $< -o $@

oracle:
	$(MAKE) -C oracle all

sass: $(LIB)
	$(CUDA_HOME)/bin/cuobjdump -sass $(LIB) > hast_b200/lib/libhast_b200.sass

clean:
	rm -rf hast_b200/lib bin
	$(MAKE) -C oracle clean
