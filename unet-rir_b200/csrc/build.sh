#!/usr/bin/env bash
# Builds liburir.so (sm_100a only) next to the Python package. Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../liburir.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
srcs=(capi.cu conv_simt.cu conv_stem.cu conv_igemm.cu conv_wgrad_tc.cu conv_thin.cu conv_head.cu conv_halo.cu conv_deep.cu conv_wgrad_halo.cu elementwise.cu vector_block.cu stft.cu)
mkdir -p "$here/build"
pids=()
for s in "${srcs[@]}"; do
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
      --expt-relaxed-constexpr "$@" -c "$here/$s" -o "$here/build/${s%.cu}.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
objs=()
for s in "${srcs[@]}"; do objs+=("$here/build/${s%.cu}.o"); done
"$NVCC" -shared -o "$out" "${objs[@]}" -lcudart_static -lrt -lpthread -ldl
echo "built $out"
