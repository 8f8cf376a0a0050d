#!/bin/bash
# Kernel A/B experiments: build ddpm_ood_b200/csrc/experiments/variants/lib_<name>.so with extra nvcc flags applied to the
# given source files (the other objects are reused from the in-tree build). Load it with DDPM_LIB_VARIANT=<path>.
#   scripts/build_variant.sh a5b5 "conv_halo.cu" -DHALO_A256=5 -DHALO_B256=5
set -e
name=$1; files=$2; shift 2
cd "$(dirname "$0")/../ddpm_ood_b200/csrc"
mkdir -p experiments/variants
objs=""
for f in *.cu; do
  o="${f%.cu}.o"
  if [[ " $files " == *" $f "* ]]; then
    o="experiments/variants/${f%.cu}_$name.o"
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c "$f" -o "$o"
  fi
  objs="$objs $o"
done
/usr/local/cuda/bin/nvcc -shared -o "experiments/variants/lib_$name.so" $objs -lcudart_static -ldl -lrt -lpthread 2>/dev/null
echo "experiments/variants/lib_$name.so"
