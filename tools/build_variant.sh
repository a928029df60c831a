#!/bin/bash
# Builds an experimental libnfa variant: tools/build_variant.sh NAME "-DMACRO ..."  ->  nfempy_b200/csrc/variants/libnfa_NAME.so
# (select it with NFA_LIB=...; only nfa_p2tet.cu is recompiled, the other objects come from the regular build)
set -e
cd "$(dirname "$0")/../nfempy_b200/csrc"
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $2 -c nfa_p2tet.cu -o variants/nfa_p2tet_$1.o 2>&1 | grep -E "error" || true
nvcc -shared -o variants/libnfa_$1.so nfa_core.o nfa_plan.o nfa_fused.o nfa_hdiv.o variants/nfa_p2tet_$1.o nfa_tensor.o nfa_multi.o nfa_p2p.o nfa_hcurl.o 2>&1 | grep -v "deprecated\|^$" || true
ls -la variants/libnfa_$1.so
