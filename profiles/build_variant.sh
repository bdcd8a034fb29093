#!/bin/bash
# builds a variant of libbioranges.so with extra nvcc flags into variants/<name>.so (objects in a scratch directory)
#   profiles/build_variant.sh <name> "<extra nvcc flags>"        then run with BRG_LIB=variants/<name>.so
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
pkg="$root/datafusion-bio-functions_b200"
tmp="$root/variants/obj_$name"
mkdir -p "$tmp"
for f in brg_api brg_sort brg_scan brg_probe brg_pairs brg_join; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O3,-fvisibility=hidden $extra -c "$pkg/csrc/$f.cu" -o "$tmp/$f.o" 2>&1 | grep -v "deprecated-gpu-targets\|offline compilation" &
done
wait
/usr/bin/g++ -std=c++17 -O2 -fPIC -fvisibility=hidden -c "$pkg/csrc/host/brg_host.cpp" -o "$tmp/brg_host.o"
/usr/local/cuda/bin/nvcc -shared -ccbin /usr/bin/g++ -o "$root/variants/$name.so" "$tmp"/*.o -lcudart_static -lpthread -ldl -lrt 2>&1 | grep -v "deprecated-gpu-targets\|offline compilation" || true
rm -rf "$tmp"
echo "built variants/$name.so"
