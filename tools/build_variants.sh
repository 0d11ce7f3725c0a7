#!/bin/bash
# Development aid: build experiment variants of the engine library.
# usage: tools/build_variants.sh name1 "-DFLAG1 -DFLAG2" name2 "-DFLAG3" ...   -> chipmunk2d_b200/lib/var_<name>.so
cd "$(dirname "$0")/.." || exit 1
while [ $# -ge 2 ]; do
	name=$1; flags=$2; shift 2
	nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -shared $flags \
		-o chipmunk2d_b200/lib/var_$name.so chipmunk2d_b200/csrc/world.cu &
done
wait
ls -la chipmunk2d_b200/lib/var_*.so
