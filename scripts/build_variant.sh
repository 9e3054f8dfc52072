#!/bin/bash
# A/B timing builds of the library: scripts/build_variant.sh <name> <extra nvcc flags...>
# Recompiles only the conv_tc_k*.cu units with the flags and links them with the objects of the normal build into
# build_ab/lib<name>.so (select it with DEPGAN_B200_LIB=build_ab/lib<name>.so).  Never shipped.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
pk="$root/dep-gan-im_b200"
out="$root/build_ab"; mkdir -p "$out/$name"
python "$pk/build.py" > /dev/null
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for u in ${UNITS:-conv_tc_k1 conv_tc_k3 conv_tc_k5}; do
  /usr/local/cuda/bin/nvcc $FL "$@" -c "$pk/csrc/$u.cu" -o "$out/$name/$u.o" &
done
wait
objs=""
for o in "$pk"/build/*.o; do
  b=$(basename "$o")
  if [ -f "$out/$name/$b" ]; then objs="$objs $out/$name/$b"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o "$out/lib$name.so" $objs -lcudart_static -ldl -lrt -lpthread
echo "$out/lib$name.so"
