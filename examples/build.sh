#!/bin/sh
# Builds the torch-free C++ host example against the in-tree library (python -m scade_b200.build first).
set -e
cd "$(dirname "$0")/.."
CUDA=${CUDA_HOME:-/usr/local/cuda}
g++ -O2 -std=c++17 examples/render_cabi.cpp -Iinclude -I"$CUDA/include" -Lscade_b200/_lib -lscade_b200 -L"$CUDA/lib64" -lcudart \
    -Wl,-rpath,'$ORIGIN/../scade_b200/_lib' -Wl,-rpath,"$CUDA/lib64" -o examples/render_cabi
echo built examples/render_cabi
