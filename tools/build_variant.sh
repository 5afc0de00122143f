#!/bin/bash
# A/B builds of libomnigs_b200.so with extra compile flags:  tools/build_variant.sh NAME "-DFLAG ..."
# -> gpurun_variants/libomnigs_b200_NAME.so (git-ignored, travels to the GPU box); load it with OMNIGS_B200_LIB.
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NAME="$1"; shift
mkdir -p "$ROOT/gpurun_variants"
make -s -j8 -C "$ROOT/omnigs-fork_b200/csrc" BUILD="build_$NAME" OUT="$ROOT/gpurun_variants/libomnigs_b200_$NAME.so" EXTRA="$*"
echo "built gpurun_variants/libomnigs_b200_$NAME.so"
