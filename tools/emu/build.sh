#!/bin/sh
# Developer tool: compile the kernels as plain C++ (CPB_EMU, see csrc/cpb_rt.h) so their
# per-thread logic can be debugged in the GPU-less build container.  NOT part of the product,
# not built by __graft_entry__.build(), never loaded by chipmunk2d_b200.
set -e
cd "$(dirname "$0")/../.."
mkdir -p tools/emu/_build
g++ -x c++ -std=c++17 -DCPB_EMU -O1 -g -fPIC -shared -ffp-contract=off -Wall -Wno-unused-function -Wno-unused-variable \
    -o tools/emu/_build/libcpb200_emu.so chipmunk2d_b200/csrc/world.cu
echo built tools/emu/_build/libcpb200_emu.so
# host layer + scene loader against the emulated engine (same sources as the product build)
gcc -O1 -g -std=gnu99 -ffp-contract=off -fopenmp -fPIC -w -DNDEBUG -shared -o tools/emu/_build/libchipmunk_b200_emu.so \
    -I include -I chipmunk2d_b200/host chipmunk2d_b200/host/*.c -Ltools/emu/_build -lcpb200_emu -Wl,-rpath,'$ORIGIN' -lm -lpthread
gcc -O1 -g -std=gnu99 -fPIC -w -shared -o tools/emu/_build/libscene_b200_emu.so -I include -I chipmunk2d_b200/scenes \
    chipmunk2d_b200/scenes/scene_io.c -Ltools/emu/_build -lchipmunk_b200_emu -Wl,-rpath,'$ORIGIN' -lm
echo built host layer against the emulator
