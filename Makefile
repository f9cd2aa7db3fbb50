# rehuel_b200 -- builds the sm_100a library that replaces librehuel.so's implicit-RK path
# (reference Makefile:2-9,49-56 builds librehuel.so from *.cpp with clang++/Armadillo; here the
# same shared-library target is produced by nvcc, with no Armadillo/LAPACK dependency).
#
#   make            -> rehuel_b200/librehuel_b200.so   (the product)
#   make oracle     -> oracle/libirk_port.so (+ oracle/_ref/librehuel_ref.so when /root/reference exists)
#   make sass       -> profiles/ SASS + ptxas listings of the hot kernel
NVCC      ?= /usr/local/cuda/bin/nvcc
# The image exports CXX=/opt/gcc/bin/g++ whose libstdc++ link is static (see oracle/Makefile);
# use the distro compiler as nvcc's host compiler so the .so links libstdc++.so.6 dynamically.
HOSTCXX   := /usr/bin/g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unknown-pragmas \
             --expt-relaxed-constexpr -Xptxas -v
SRC       := rehuel_b200/csrc
LIB       := rehuel_b200/librehuel_b200.so
INST      := robertson van_der_pol exponential stiff_eq brusselator lorenz harmonic dimer kinetic4 three_body \
             bruss1d_64 bruss1d_32 bruss1d_16 bruss1d_12 bruss1d_8 bruss1d_dense_64 bruss1d_dense_32 bruss1d_dense_16 bruss1d_dense_12 bruss1d_dense_8
OBJS      := build/capi.o build/order.o build/selftest.o build/tableau.o $(INST:%=build/inst_%.o)
KHDRS     := $(SRC)/irk_thread.cuh $(SRC)/irk_cta.cuh $(SRC)/irk_warp.cuh $(SRC)/irk_wdense.cuh $(SRC)/irk_wseg.cuh $(SRC)/band_twisted.cuh $(SRC)/launch.cuh $(SRC)/functors.cuh \
             $(SRC)/tableau.hpp include/rehuel_b200.h

.PHONY: all oracle clean sass examples tableaux microbench
all: $(LIB)
microbench: build/microbench

build:
	mkdir -p build

build/capi.o: $(SRC)/capi.cu $(KHDRS) | build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/capi.ptxas.log || (cat build/capi.ptxas.log; false)

# one translation unit per built-in functor: `make -j` compiles the kernel instantiations in parallel
build/inst_%.o: $(SRC)/instances/inst_%.cu $(KHDRS) | build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/inst_$*.ptxas.log || (cat build/inst_$*.ptxas.log; false)

build/order.o: $(SRC)/order.cu $(SRC)/launch.cuh include/rehuel_b200.h | build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/order.ptxas.log || (cat build/order.ptxas.log; false)

build/selftest.o: $(SRC)/selftest.cu $(KHDRS) | build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> build/selftest.ptxas.log || (cat build/selftest.ptxas.log; false)

build/tableau.o: $(SRC)/tableau.cpp $(SRC)/tableau.hpp $(SRC)/tableau_eig.inc $(SRC)/tableau_dec.inc | build
	$(HOSTCXX) -O2 -std=c++17 -fPIC -Wall -Wextra -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $(OBJS) -cudart shared

# Regenerates the committed tableau_eig.inc / tableau_dec.inc (+ oracle/tableau_dec.inc); needs mpmath.
# Deliberately not a file dependency: a fresh checkout must build without running the generator.
tableaux:
	python tools/gen_tableaux.py

# Programs on the three faces of the boundary: C ABI, C++ API (built-in problem), C++ user functor.
examples: build/c_interface_test build/cpp_api build/user_functor build/rehuel_example
build/c_interface_test: examples/c_interface_test.c include/rehuel_b200.h $(LIB)
	/usr/bin/gcc -std=c11 -O2 -Wall -Iinclude $< -o $@ -Lrehuel_b200 -lrehuel_b200 -lm -Wl,-rpath,'$$ORIGIN/../rehuel_b200'
build/cpp_api: examples/cpp_api.cpp include/rehuel_b200/irk.hpp $(LIB)
	$(HOSTCXX) -std=c++11 -O2 -Wall -Iinclude $< -o $@ -Lrehuel_b200 -lrehuel_b200 -Wl,-rpath,'$$ORIGIN/../rehuel_b200'
build/rehuel_example: examples/rehuel_example.cpp include/rehuel_b200/irk.hpp include/rehuel_b200.h $(LIB)
	$(HOSTCXX) -std=c++11 -O2 -Wall -Iinclude $< -o $@ -Lrehuel_b200 -lrehuel_b200 -Wl,-rpath,'$$ORIGIN/../rehuel_b200'
build/user_functor: examples/user_functor.cu include/rehuel_b200/irk.cuh include/rehuel_b200/irk.hpp $(KHDRS) $(LIB)
	$(NVCC) -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) --expt-relaxed-constexpr -Iinclude $< -o $@ \
	    -Lrehuel_b200 -lrehuel_b200 -cudart shared -Xlinker -rpath,'$$ORIGIN/../rehuel_b200'

# stand-alone micro-benchmarks behind DESIGN.md 5.1 (issue model) and 5.5 (DMMA vs FMA pipe)
build/microbench: tools/microbench.cu | build
	$(NVCC) -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) $< -o $@

oracle:
	$(MAKE) -C oracle all

sass: $(LIB)
	mkdir -p profiles
	/usr/local/cuda/bin/cuobjdump -sass $(LIB) > profiles/librehuel_b200.sass
	cat build/inst_*.ptxas.log > profiles/ptxas_v.log

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
