#!/bin/sh
# tests/emul/build.sh — builds the host lock-step emulator (test infrastructure only).
set -e
cd "$(dirname "$0")"
mkdir -p _build
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
DEFS=""
[ -f ../../directxtex_b200/csrc/dxb_bc7.cuh ] && DEFS="-DDXB_EMUL_BC7"
$CXX -std=c++17 -O2 -msse2 -mfpmath=sse -mfma -ffp-contract=off -fopenmp -fPIC -shared -x c++ $DEFS \
    -I../../directxtex_b200/csrc emul.cpp -o _build/libdxb_emul.so
echo built tests/emul/_build/libdxb_emul.so
