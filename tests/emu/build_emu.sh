#!/bin/bash
# Builds the kernels of free_hunch_b200/csrc for the CPU emulation shim (tools/cuda_emu.h) so that the
# index/logic of every kernel can be exercised without a GPU.  TEST INFRASTRUCTURE: the product library is
# built by nvcc (see __graft_entry__.build).
set -e
cd "$(dirname "$0")/../.."
g++ -O2 -g -std=c++17 -fPIC -shared -pthread -DFH_EMU -include tools/cuda_emu.h -x c++ \
    free_hunch_b200/csrc/fh_api.cu tests/emu/emu_globals.cpp -o tests/emu/libfreehunch_emu.so
echo built tests/emu/libfreehunch_emu.so
