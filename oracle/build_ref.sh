#!/bin/sh
# Compile the REFERENCE's own CUDA implementation of the path from its sources where they lie under
# /root/reference (read-only) into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun).
# Recipe = the reference's setup.py:28-48,64-77 file list and flags plus an sm_100 arch flag; its own build
# system is not run and no reference source is copied into the repository.
# The reference's Python wrapper package is installed next to it (like `pip install --target`) so that the
# "unchanged pyparament wrapper on our library" acceptance tests can also run on the GPU box.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
[ -d "$REF/src/cuda" ] || { echo "reference tree not found at $REF: keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT"
cd "$REF/src/cuda"
nvcc -lcublas -DPARAMENT_BUILD_DLL -DNDEBUG --shared --compiler-options -fPIC \
     -gencode arch=compute_100,code=sm_100 -o "$OUT/libparament.so" \
     deviceInfo.c diagonal_add.cu mathhelper.cpp parament.cpp printFuncs.cpp debugfuncs.cpp control_expansion.cu
rm -rf "$OUT/pyparament"
mkdir -p "$OUT/pyparament"
cp -r "$REF/src/python/pyparament/parament" "$OUT/pyparament/parament"
rm -rf "$OUT/pyparament/parament/__pycache__" "$OUT/pyparament/parament/test/__pycache__"
echo "built $OUT/libparament.so and installed the reference wrapper into $OUT/pyparament"
