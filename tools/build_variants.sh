#!/bin/bash
# Build A/B variants of the variance GEMM (dev tool): tools/build_variants.sh "NAME:-DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/../bosip.jl_b200"
python build.py > /dev/null
mkdir -p csrc/build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c csrc/vargemm.cu -o csrc/build/variants/vargemm_$name.o
  objs=$(ls csrc/build/*.o | grep -v vargemm.o)
  nvcc -shared -o csrc/build/variants/libbosip_b200_$name.so $objs csrc/build/variants/vargemm_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
  echo built $name
done
