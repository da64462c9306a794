#!/usr/bin/env bash
# Builds libmetaasr_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libmetaasr_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "$HERE/build"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  stale=0
  for h in "$HERE"/*.cuh "$HERE/../../include/metaasr_b200.h"; do [[ "$h" -nt "$o" ]] && stale=1; done
  if [[ ! -f "$o" || "$f" -nt "$o" || $stale == 1 ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/build/*.o -cudart static
echo "built $OUT"
