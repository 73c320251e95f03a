#!/bin/bash
# builds profiles/debug/libplume_b200_<name>.so with extra nvcc flags: build_variant.sh <name> <flags...>
# (A/B measurements of build-time knobs: python profiles/debug/variant_bench.py profiles/debug/libplume_b200_<name>.so)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/../.." && pwd)
out=$root/profiles/debug/libplume_b200_$name.so
tmp=$(mktemp -d)
for f in $root/uav-wrf-les-ppo-lstm_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $tmp/$(basename $f .cu).o &
done
wait
for f in $root/uav-wrf-les-ppo-lstm_b200/csrc/*.cu; do test -f $tmp/$(basename $f .cu).o || { echo "compile of $f failed"; exit 1; }; done
nvcc -shared -o $out $tmp/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
rm -rf $tmp
echo $out
