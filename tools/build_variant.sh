#!/bin/bash
# Dev aid: build a library variant with extra -D flags:  tools/build_variant.sh NAME [-DREHUEL_MINB=5 ...]
# -> build/var/NAME.so; time it on the GPU with tools/gpu_dev.py variants build/var/NAME.so
# (only the Robertson and Van der Pol instantiations are rebuilt with the flags; the rest is linked as built)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var
make -j8 >/dev/null
for inst in robertson van_der_pol; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v "$@" -c rehuel_b200/csrc/instances/inst_$inst.cu -o build/var/$name.$inst.o 2> build/var/$name.$inst.ptxas.log &
done
wait
others=$(ls build/*.o | grep -v "inst_robertson.o\|inst_van_der_pol.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o build/var/$name.so build/var/$name.robertson.o build/var/$name.van_der_pol.o $others -cudart shared
rm -f build/var/$name.*.o
grep -A2 "RobertsonFELi3ELi1ELi1ELi1E.*Lb0ELb0" build/var/$name.robertson.ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
