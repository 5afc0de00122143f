#!/usr/bin/env bash
# Build oracle/_ref/omnigs_ref.so: the reference's own rasterizer, sources compiled WHERE THEY LIE
# under /root/reference (nothing is copied into this repo), for sm_100, default -fmad, no fast-math
# (the reference's Release flags, CMakeLists.txt:4-7).  Only additions: our glm-compat shim on the
# include path (glm is an absent third-party dependency) and a force-included fix-up header.
# Outputs go to oracle/_ref/ only (git-ignored, shipped to the GPU box by gpurun).
# The reference's own build system (cmake + OpenCV/Eigen/jsoncpp/...) is NOT run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${OGS_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -d "$REF/cuda_rasterizer" ]; then
	echo "build_ref.sh: $REF not present (GPU box?) - using prebuilt $OUT/omnigs_ref.so if any" >&2
	exit 0
fi
mkdir -p "$OUT"
TORCH="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))' 2>/dev/null)"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC
       -include "$HERE/shim/fix.h" -I"$HERE/shim" -I"$REF" -I"$REF/cuda_rasterizer"
       -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" -I"$PYINC"
       -D_GLIBCXX_USE_CXX11_ABI=1 --expt-relaxed-constexpr -w)
pids=()
build() { # src obj
	if [ ! -f "$2" ] || [ "$1" -nt "$2" ] || [ "$HERE/shim/glm/glm.hpp" -nt "$2" ] || [ "$HERE/shim/fix.h" -nt "$2" ]; then
		echo "  nvcc $1"
		"$NVCC" "${FLAGS[@]}" -c "$1" -o "$2" &
		pids+=($!)
	fi
}
build "$REF/cuda_rasterizer/forward.cu"         "$OUT/forward.o"
build "$REF/cuda_rasterizer/backward.cu"        "$OUT/backward.o"
build "$REF/cuda_rasterizer/rasterizer_impl.cu" "$OUT/rasterizer_impl.o"
build "$REF/src/rasterize_points.cu"            "$OUT/rasterize_points.o"
build "$HERE/ref_module.cu"                     "$OUT/ref_module.o"
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/omnigs_ref.so" "$OUT"/forward.o "$OUT"/backward.o "$OUT"/rasterizer_impl.o \
	"$OUT"/rasterize_points.o "$OUT"/ref_module.o \
	-Xlinker -Bsymbolic -Xlinker -rpath -Xlinker "$TORCH/lib" \
	-L"$TORCH/lib" -ltorch -ltorch_cpu -ltorch_cuda -ltorch_python -lc10 -lc10_cuda -lcudart
echo "built $OUT/omnigs_ref.so"
