#!/bin/bash
# Dev aid: build a library variant with extra -D flags:  tools/build_variant.sh NAME [-DREHUEL_MINB=5 ...]
# -> build/var/NAME.so (+ NAME.ptxas.log); time it on the GPU with tools/time_variants.py build/var/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var
make build/tableau.o build/order.o >/dev/null
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c rehuel_b200/csrc/capi.cu -o build/var/$name.o 2> build/var/$name.ptxas.log
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/var/$name.so build/var/$name.o build/order.o build/tableau.o -cudart shared
rm -f build/var/$name.o
grep -A1 "ensemble_irk_threadINS_10RobertsonFELi3ELi1ELi1ELi1E.*Lb0" build/var/$name.ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
