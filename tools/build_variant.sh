#!/bin/sh
# Build the CUDA library with extra compile-time knobs into variants/<name>/libagbnp_b200.so (git-ignored; travels to the
# GPU box):   tools/build_variant.sh chunk16 -DGB_CHUNK_TILES=16
# and time it with   AGBNP_B200_LIB=variants/chunk16/libagbnp_b200.so python tools/quick_time.py hivrt
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/variants/$name"
cd "$root/openmm_agbnp_plugin_b200"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
     -o "$root/variants/$name/libagbnp_b200.so" csrc/agbnp_b200.cu csrc/agbnp_setup.cpp
