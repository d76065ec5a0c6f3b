# acx-b200 build: libacx.so (runtime + bundled mini-MPI), acxrun launcher, C/CUDA tests.
# The reference builds a static lib with nvcc -dc for sm_80 (reference Makefile:3,8,36);
# we target sm_100a only and need no relocatable device code (device API is header-inline).
NVCC      ?= nvcc
CXX       ?= g++
GENCODE   ?= -gencode arch=compute_100a,code=sm_100a
# the reference's spelling of the same knob (reference Makefile:3, README.md:75-81); anything but sm_100a will not compile (tcgen05 / TMA PTX)
ifdef NVCC_GENCODE
GENCODE := $(NVCC_GENCODE)
endif
NVCCFLAGS ?= $(GENCODE) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unused-parameter -Iinclude
CXXFLAGS  ?= -O2 -std=c++17 -fPIC -Wall -Wextra -Iinclude
B := build

LIB   := lib/libacx.so
MPILIB := lib/libacxmpi.a
BINS  := bin/acxrun
BENCH := $(patsubst bench/%.cu,bin/%,$(wildcard bench/*.cu))
CTESTS := $(patsubst tests/c/%.c,$(B)/tests/%,$(wildcard tests/c/*.c)) $(patsubst tests/c/%.cu,$(B)/tests/%,$(wildcard tests/c/*.cu))

.PHONY: all clean lib tests bench examples tsan asan install extmpi-check extmpi-lib emu emu-tsan emu-lib emu-tests emu-asan emu-tsan-matrix emu-reftests
all: lib $(BINS) tests bench
bench: $(BENCH)
bin/%: bench/%.cu $(LIB)
	@mkdir -p bin
	$(NVCC) $(GENCODE) -O3 -std=c++17 -lineinfo -Iinclude -Ibench $< -o $@ -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../lib' -cudart static
lib: $(LIB) $(MPILIB)

$(B)/minimpi.o: src/minimpi/minimpi.cpp include/mpi.h
	@mkdir -p $(B)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(MPILIB): $(B)/minimpi.o
	@mkdir -p lib
	ar rcs $@ $^

$(B)/acx.o: src/core/acx.cu src/core/internal.h src/core/kernels.cuh include/mpi-acx.h include/acx_device.cuh include/mpi.h
	@mkdir -p $(B)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

KOBJS := $(patsubst src/kernels/%.cu,$(B)/k_%.o,$(wildcard src/kernels/*.cu))
$(B)/k_%.o: src/kernels/%.cu include/mpi-acx.h include/acx_device.cuh include/acx_kernels.h src/kernels/gemm_fp8_sm100.cuh
	@mkdir -p $(B)
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> $(B)/k_$*.ptxas.log || (cat $(B)/k_$*.ptxas.log; false)

$(LIB): $(B)/acx.o $(B)/minimpi.o $(KOBJS) src/core/acx.map
	@mkdir -p lib
	$(NVCC) $(GENCODE) -shared -o $@ $(filter %.o,$^) -cudart static -ldl -lpthread -lrt -Xlinker --version-script=src/core/acx.map
	ln -sf libacx.so lib/libmpi-acx.so   # the reference's library name: `-lmpi-acx` keeps working

bin/acxrun: launcher/acxrun.cpp
	@mkdir -p bin
	$(CXX) -O2 -std=c++17 -Wall $< -o $@

tests: $(CTESTS)
$(B)/tests/%: tests/c/%.c $(LIB)
	@mkdir -p $(B)/tests
	$(NVCC) $(GENCODE) -O2 -Iinclude $< -o $@ -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../../lib' -cudart static
$(B)/tests/%: tests/c/%.cu $(LIB)
	@mkdir -p $(B)/tests
	$(NVCC) $(GENCODE) -O2 -lineinfo -Iinclude $< -o $@ -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../../lib' -cudart static

examples: $(LIB)
	@mkdir -p $(B)/examples
	$(NVCC) $(GENCODE) -O2 -Iinclude examples/stream_ring.c -o $(B)/examples/stream_ring -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../../lib' -cudart static
	$(NVCC) $(GENCODE) -O2 -Iinclude examples/token_ring.c -o $(B)/examples/token_ring -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../../lib' -cudart static
	$(NVCC) $(GENCODE) -O2 -Iinclude examples/partitioned_producer.cu -o $(B)/examples/partitioned_producer -Llib -lacx -Xlinker -rpath,'$$ORIGIN/../../lib' -cudart static

# Portability check (compile only, no MPI needed): build the runtime and the fused kernels against a declarations-only
# MPI whose handles are pointers to opaque structs (Open MPI style) instead of the bundled mini-MPI's integers. Proves
# nothing outside the #ifdef ACX_MINI_MPI sections depends on the bundled library.
extmpi-check:
	@mkdir -p $(B)/extmpi
	for f in src/core/acx.cu src/kernels/halo_stencil.cu src/kernels/pp_gemm.cu; do \
	  $(NVCC) $(GENCODE) -std=c++17 -Itools/extmpi_stub -Iinclude -c $$f -o $(B)/extmpi/check.o || exit 1; done
	@if nm -u $(B)/extmpi/check.o | grep -q ACX_Mpi; then echo "bundled-MPI extension used outside its guard"; exit 1; fi
	@echo "extmpi-check: OK"

# Build against an external MPI instead of the bundled one:  make extmpi-lib MPI_HOME=/opt/openmpi
# (needs MPI-4 partitioned calls only for the host transport; produces lib/libacx_extmpi.so; untested here: the image has no MPI)
extmpi-lib: $(KOBJS)
	@test -n "$(MPI_HOME)" || (echo "set MPI_HOME"; exit 1)
	@mkdir -p lib $(B)/extmpi
	$(NVCC) $(GENCODE) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$(MPI_HOME)/include -Iinclude -c src/core/acx.cu -o $(B)/extmpi/acx.o
	for f in halo_stencil pp_gemm; do $(NVCC) $(GENCODE) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$(MPI_HOME)/include -Iinclude -c src/kernels/$$f.cu -o $(B)/extmpi/k_$$f.o || exit 1; done
	$(NVCC) $(GENCODE) -shared -o lib/libacx_extmpi.so $(B)/extmpi/acx.o $(B)/extmpi/k_halo_stencil.o $(B)/extmpi/k_pp_gemm.o -cudart static -L$(MPI_HOME)/lib -lmpi -ldl -lpthread -lrt

# Install headers, the library and the launcher (PREFIX=/usr/local by default): a program written for the reference then
# builds with  nvcc -I$(PREFIX)/include prog.cu -L$(PREFIX)/lib -lacx
PREFIX ?= /usr/local
install: lib bin/acxrun
	install -d $(DESTDIR)$(PREFIX)/include $(DESTDIR)$(PREFIX)/lib $(DESTDIR)$(PREFIX)/bin
	install -m 644 include/mpi-acx.h include/acx_device.cuh include/acx_kernels.h include/mpi.h $(DESTDIR)$(PREFIX)/include/
	install -m 755 lib/libacx.so $(DESTDIR)$(PREFIX)/lib/
	ln -sf libacx.so $(DESTDIR)$(PREFIX)/lib/libmpi-acx.so
	install -m 644 lib/libacxmpi.a $(DESTDIR)$(PREFIX)/lib/
	install -m 755 bin/acxrun $(DESTDIR)$(PREFIX)/bin/

# ---- no-GPU emulation (tests/emu) ---------------------------------------------------------------------------------
# `make emu`: the protocol kernels of src/core/kernels.cuh compiled as plain C++ and called directly, several ranks as
# threads of one process (one OS thread per CTA, one fiber per CUDA thread). No GPU, no libacx.so.
# `make emu-lib emu-tests`: build/emu/libacx_emu.so = the UNMODIFIED runtime (src/core/acx.cu as C++) + a software CUDA
# runtime (tests/emu/cudart_emu.cpp: streams, events, capture, graphs, IPC memory) + the mini-MPI; the conformance
# programs link against it and run under acxrun, one process per rank, on the *device* transport.
CUDA_HOME ?= /usr/local/cuda
SAN_CXX ?= $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
EMU := $(B)/emu
EMU_FLAGS := -std=c++17 -g -I$(CUDA_HOME)/include -Iinclude -Isrc/core -Itests/emu
EMU_SAN ?=
emu:
	@mkdir -p $(B)/tests
	$(CXX) -O1 $(EMU_FLAGS) tests/emu/protocol_emu.cpp -o $(B)/tests/protocol_emu -lpthread
	$(B)/tests/protocol_emu

# The same program under ThreadSanitizer (ranks and CTAs are threads of one process, so cross-rank accesses are visible to it):
# checks the release/acquire pairing of the flag protocol against the C++ memory model.
emu-tsan:
	@mkdir -p $(B)/tests
	$(SAN_CXX) -O1 -fsanitize=thread -Wno-tsan $(EMU_FLAGS) tests/emu/protocol_emu.cpp -o $(B)/tests/protocol_emu_tsan -lpthread
	TSAN_OPTIONS="halt_on_error=1 history_size=4" $(B)/tests/protocol_emu_tsan

EMU_CTESTS := $(patsubst tests/c/%.c,$(EMU)/tests/%,$(wildcard tests/c/*.c)) $(EMU)/tests/ring_partitioned $(EMU)/tests/partitioned_device
emu-lib: $(EMU)/libacx_emu.so
emu-tests: $(EMU_CTESTS) $(EMU)/p2p_bench
$(EMU)/acx.o: src/core/acx.cu src/core/internal.h src/core/kernels.cuh include/mpi-acx.h include/acx_device.cuh include/mpi.h tests/emu/cuda_emu.h
	@mkdir -p $(EMU)
	$(CXX) -O1 -fPIC $(EMU_FLAGS) $(EMU_SAN) -x c++ -include tests/emu/cuda_emu.h -c $< -o $@
$(EMU)/cudart_emu.o: tests/emu/cudart_emu.cpp tests/emu/cuda_emu.h
	@mkdir -p $(EMU)
	$(CXX) -O1 -fPIC -Wall $(EMU_FLAGS) $(EMU_SAN) -c $< -o $@
$(EMU)/minimpi.o: src/minimpi/minimpi.cpp include/mpi.h
	@mkdir -p $(EMU)
	$(CXX) $(CXXFLAGS) -g $(EMU_SAN) -c $< -o $@
$(EMU)/libacx_emu.so: $(EMU)/acx.o $(EMU)/cudart_emu.o $(EMU)/minimpi.o
	$(CXX) -shared $(EMU_SAN) -Wl,-Bsymbolic-functions -o $@ $^ -ldl -lpthread -lrt   # own cuda* / MPI* definitions win even inside a process that also loaded the real libcudart (python + torch)
$(EMU)/tests/%: tests/c/%.c $(EMU)/libacx_emu.so tests/c/acx_test.h
	@mkdir -p $(EMU)/tests
	gcc -O1 -g $(EMU_SAN) -Iinclude -I$(CUDA_HOME)/include $< -o $@ -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU)
# Drop-in check: the REFERENCE's own test sources (read in place from $(REF), nothing is copied), unmodified, built against
# include/ + the software device. The two graph tests use CUDA 11's five-argument cudaGraphInstantiate, which only the C++
# runtime header still offers, so they are compiled as C++; ring-partitioned.cu is piped through sed, which spells its two
# kernel<<<...>>>(...) launches as function calls (kernels are plain functions on the software device).
REF ?= /root/reference
emu-reftests: $(EMU)/libacx_emu.so
	@mkdir -p $(EMU)/reftests
	for t in ring ring-all ring-all-device; do gcc -O1 -DDEBUG -Iinclude -I$(CUDA_HOME)/include $(REF)/test/src/$$t.c -o $(EMU)/reftests/$$t -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU) || exit 1; done
	for t in ring-all-graph ring-all-graph-construction; do $(CXX) -x c++ -O1 -DDEBUG -fpermissive -w -Iinclude -I$(CUDA_HOME)/include $(REF)/test/src/$$t.c -o $(EMU)/reftests/$$t -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU) || exit 1; done
	sed -E 's/([A-Za-z_][A-Za-z_0-9]*)<<<([^,>]*),([^,>]*),([^,>]*),([^,>]*)>>>\((.*)\);/emu::launch_user(\1, \2, \3, (cudaStream_t)(\5), \6);/' $(REF)/test/src/ring-partitioned.cu | \
	  $(CXX) -x c++ -O1 -DDEBUG -fpermissive -w $(EMU_FLAGS) -include tests/emu/cuda_emu.h - -o $(EMU)/reftests/ring-partitioned -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU) -lpthread   # kernel<<<...>>>(...) spelled as a function call on the fly

$(EMU)/p2p_bench: bench/p2p_bench.cu $(EMU)/libacx_emu.so tests/emu/cuda_emu.h
	$(CXX) -O1 $(EMU_FLAGS) $(EMU_SAN) -x c++ -include tests/emu/cuda_emu.h $< -o $@ -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU) -lpthread
$(EMU)/tests/%: tests/c/%.cu $(EMU)/libacx_emu.so tests/c/acx_test.h tests/emu/cuda_emu.h
	@mkdir -p $(EMU)/tests
	$(CXX) -O1 $(EMU_FLAGS) $(EMU_SAN) -x c++ -include tests/emu/cuda_emu.h $< -o $@ -L$(EMU) -lacx_emu -Wl,-rpath,$(CURDIR)/$(EMU) -lpthread

# ASan + UBSan (+ LeakSanitizer) build of the software-device stack and a run of the whole matrix on it
emu-asan:
	$(MAKE) EMU=$(B)/emu_asan EMU_SAN="-fsanitize=address,undefined -fno-omit-frame-pointer" CXX=$(SAN_CXX) emu-lib emu-tests
	ASAN_OPTIONS=detect_leaks=1:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1 EMU_DIR=$(B)/emu_asan tools/emu_matrix.sh quick

# ThreadSanitizer build of the software-device stack (host runtime, stream workers, CTAs of one process) + matrix
emu-tsan-matrix:
	$(MAKE) EMU=$(B)/emu_tsan EMU_SAN="-fsanitize=thread -Wno-tsan" CXX=$(SAN_CXX) emu-lib emu-tests
	TSAN_OPTIONS="halt_on_error=0 history_size=4" EMU_DIR=$(B)/emu_tsan tools/emu_matrix.sh tsan

clean:
	rm -rf $(B) lib bin

# ThreadSanitizer build of the mini-MPI + its threaded test (host-side race detection; the reference has none)
tsan:
	@mkdir -p $(B)
	$(SAN_CXX) -O1 -g -fsanitize=thread -std=c++17 -Iinclude -c src/minimpi/minimpi.cpp -o $(B)/minimpi_tsan.o
	gcc -O1 -g -fsanitize=thread -Iinclude -c tests/c/minimpi_threads.c -o $(B)/mt_tsan.o
	$(SAN_CXX) -fsanitize=thread $(B)/mt_tsan.o $(B)/minimpi_tsan.o -o $(B)/tests/minimpi_threads_tsan -ldl -lpthread -lrt
	bin/acxrun -np 2 --timeout 100 $(B)/tests/minimpi_threads_tsan

# AddressSanitizer + UBSan build of the host side (mini-MPI + runtime front-ends on the host transport)
asan:
	@mkdir -p $(B)/asan
	$(SAN_CXX) -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -Iinclude -c src/minimpi/minimpi.cpp -o $(B)/asan/minimpi.o
	$(NVCC) $(GENCODE) -O1 -std=c++17 -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -Iinclude -c src/core/acx.cu -o $(B)/asan/acx.o
	$(NVCC) $(GENCODE) -shared -o $(B)/asan/libacx_asan.so $(B)/asan/acx.o $(B)/asan/minimpi.o -cudart static -ldl -lpthread -lrt -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined
	for t in ring ring_all minimpi_smoke minimpi_comms minimpi_types procnull_comms wildcards subcomm datatypes; do gcc -O1 -g -fsanitize=address,undefined -Iinclude -I/usr/local/cuda/include tests/c/$$t.c -o $(B)/asan/$$t -L$(B)/asan -lacx_asan -Wl,-rpath,$(CURDIR)/$(B)/asan -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,/usr/local/cuda/lib64 && ASAN_OPTIONS=detect_leaks=0 ACX_FORCE_HOST=1 bin/acxrun -np 3 --timeout 100 $(B)/asan/$$t || exit 1; done
