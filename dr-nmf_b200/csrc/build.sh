#!/bin/bash
# Build libdrnmf.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libdrnmf.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
mkdir -p "$HERE/../build"
OBJS=""
for f in runtime prep gemm stft recurrent_simt recurrent_tc snmf train api; do
  src="$HERE/$f.cu"; obj="$HERE/../build/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/common.cuh" -nt "$obj" ] || [ "$HERE/internal.h" -nt "$obj" ] || [ "$HERE/gemm_simt.cuh" -nt "$obj" ] || [ "$HERE/../../include/drnmf.h" -nt "$obj" ]; then
    ( $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$src" -o "$obj" || { rm -f "$obj"; echo "FAILED: $f.cu" >&2; } ) &
  fi
  OBJS="$OBJS $obj"
done
wait
for o in $OBJS; do [ -f "$o" ] || { echo "build failed: $o missing" >&2; exit 1; }; done
$NVCC -shared -o "$OUT" $OBJS -gencode arch=compute_100a,code=sm_100a -cudart static
echo "built $OUT"
