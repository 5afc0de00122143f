#!/usr/bin/env bash
# Builds omnigs-fork_b200/omnigs_b200_torch.so: the LibTorch drop-in (rasterize_points.cpp) plus a
# pybind11 module, linked against libomnigs_b200.so (rpath $ORIGIN).  Host compiler only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
PKG="$(dirname "$HERE")"
PY="${PYTHON:-python}"
OUT="$PKG/omnigs_b200_torch.so"
newest=$(ls -t "$HERE"/rasterize_points.cpp "$HERE"/rasterize_points.h "$HERE"/torch_module.cpp "$PKG/../include/omnigs_b200.h" | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ]; then exit 0; fi
TORCH="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null)"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
mkdir -p "$HERE/build"
FLAGS=(-std=c++17 -O2 -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include"
       -I"$PYINC" -I/usr/local/cuda/include -w)
g++ "${FLAGS[@]}" -c "$HERE/rasterize_points.cpp" -o "$HERE/build/rasterize_points.o" &
g++ "${FLAGS[@]}" -c "$HERE/torch_module.cpp" -o "$HERE/build/torch_module.o" &
wait
g++ -shared -o "$OUT" "$HERE/build/rasterize_points.o" "$HERE/build/torch_module.o" \
	-L"$PKG" -lomnigs_b200 -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$TORCH/lib" \
	-L"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -ltorch_python -lc10 -lc10_cuda -L/usr/local/cuda/lib64 -lcudart
echo "built $OUT"
