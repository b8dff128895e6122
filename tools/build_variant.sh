#!/bin/bash
# A/B experiments: build the library with extra nvcc flags into tfmq-dm_b200/tfmq_b200/lib/<name>.so
#   tools/build_variant.sh libA.so -DTFMQ_EPI_WARPS=8
# (on the GPU box: cp lib/<name>.so lib/libtfmq_b200.so before the run to be measured)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/tfmq-dm_b200/csrc
obj=$root/tfmq-dm_b200/build/variant_${name%.so}
mkdir -p "$obj" "$root/tfmq-dm_b200/tfmq_b200/lib"
for f in "$src"/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -I "$root/include" "$@" -c "$f" -o "$obj/$(basename "${f%.cu}").o" &
done
wait
nvcc -shared -o "$root/tfmq-dm_b200/tfmq_b200/lib/$name" "$obj"/*.o -cudart static -gencode arch=compute_100a,code=sm_100a
echo "built $name"
