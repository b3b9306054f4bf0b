#!/bin/bash
# Builds the CPU thread emulator of the CUDA kernels (test infrastructure, see emu.cpp).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
g++ -std=c++20 -O2 -ffp-contract=off -fPIC -shared -DB2R_HOST_EMU -pthread \
    -I/usr/local/cuda/include \
    "$here/emu.cpp" "$here/../../vkresample_b200/csrc/b2r_plan.cpp" \
    -o "$here/libb2r_emu.so"
