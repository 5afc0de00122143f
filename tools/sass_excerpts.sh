#!/bin/bash
# SASS of the kernels the design claims rest on (blend loops, onesweep pass), from the in-tree objects:
#   tools/sass_excerpts.sh profiles/r02_sass      (no GPU needed)
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
OUT="${1:-$ROOT/profiles/sass}"; mkdir -p "$OUT"
dump() { # object, function-name regex, output file
	cuobjdump -sass "$ROOT/omnigs-fork_b200/csrc/build/$1" 2>/dev/null |
	awk -v pat="$2" '/Function :/{f = ($0 ~ pat)} f' | grep -E "Function :|^\s+/\*[0-9a-f]{4}\*/" |
	sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/[[:space:]]+$//' > "$OUT/$3"
	echo "$3: $(grep -c '/\*' "$OUT/$3") instructions"
}
dump render_fwd.o 'render_fwd_kernelE|render_fwd_list_kernel' render_fwd.sass
dump render_bwd.o 'render_bwd_kernelILi16' render_bwd.sass
dump binning.o 'onesweep_pass_kernelILi6ELi3|onesweep_pass_kernelILi7ELi2|onesweep_pass_kernelILi8ELi0' onesweep_pass.sass
dump preprocess_fwd.o 'preprocess_lonlat_fwd_kernelILi1ELb0' preprocess_fwd.sass
