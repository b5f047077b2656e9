#!/usr/bin/env bash
# Build libtt_b200.so (sm_100a only) next to this script.  Used by
# __graft_entry__.build() and by hand:  bash recommendsystemproject_b200/csrc/build.sh
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
pids=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ tc_common.cuh -nt "$o" ] || [ topk_common.cuh -nt "$o" ] || [ ../../include/tt_b200.h -nt "$o" ]; then
    ( $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -c "$f" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
$NVCC -shared -o libtt_b200.so build/*.o
echo "built $(pwd)/libtt_b200.so"
