#!/bin/bash
# Builds vnlb_b200/libvnlb_b200.so (sm_100a) and the CPU oracle.  Every .cu is its own translation unit (no device
# symbols cross files), compiled in parallel into build/ and linked into one shared library.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA}"
mkdir -p build
pids=()
objs=()
for src in vnlb_b200/csrc/*.cu; do
    obj=build/$(basename "${src%.cu}").o
    objs+=("$obj")
    if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ -n "$(find vnlb_b200/csrc -name '*.cuh' -newer "$obj")" ] || [ include/vnlb_b200.h -nt "$obj" ] || [ -n "$NVCC_EXTRA" ]; then
        $NVCC $FLAGS -c -o "$obj" "$src" &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o vnlb_b200/libvnlb_b200.so "${objs[@]}"
make -s -C oracle
