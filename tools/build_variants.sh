#!/bin/bash
# Build A/B variants of one kernel file (dev tool): tools/build_variants.sh FILE.cu "NAME:-DFLAG=.. -DFLAG=.." ...
# -> bosip.jl_b200/csrc/build/variants/libbosip_b200_NAME.so (select with BOSIP_B200_LIB=...)
set -e
cd "$(dirname "$0")/../bosip.jl_b200"
python build.py > /dev/null
mkdir -p csrc/build/variants
file="$1"; shift
base="${file%.cu}"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c csrc/$file -o csrc/build/variants/${base}_$name.o
  objs=$(ls csrc/build/*.o | grep -v "/$base.o" | grep -v "/${base}_k[0-9].o")
  nvcc -shared -o csrc/build/variants/libbosip_b200_$name.so $objs csrc/build/variants/${base}_$name.o -gencode arch=compute_100a,code=sm_100a -lcudart
  echo built $name
done
