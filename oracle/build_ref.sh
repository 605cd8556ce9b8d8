#!/usr/bin/env bash
# oracle/build_ref.sh -- compile the REFERENCE'S OWN sources, where they lie,
# into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# Outputs: _ref/libkc_ref.so, _ref/ref_count.  Nothing from /root/reference is
# copied into the repository: the generated kernel include lives in a temp dir
# for the duration of the build only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${KC_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/GPUHandler.cu" ]; then
    echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
# kernels + comparators (GPUHandler.cu:10-298) and CheckEquals/reduceKMers (:329-360);
# the only edit is the CUDA thread index -> KC_TID (one host call per read)
sed -n '10,298p;329,360p' "$REF/GPUHandler.cu" \
  | sed 's/(blockIdx.x \* blockDim.x + threadIdx.x)/KC_TID/g' > "$TMP/ref_kernels.inc"
for sym in 'void bitEncode' 'void extractKMers' 'calculateOutputSize' 'class KMer128Comparator' 'reduceKMers'; do
    grep -q "$sym" "$TMP/ref_kernels.inc" || { echo "build_ref: '$sym' not found at the expected lines" >&2; exit 1; }
done
if grep -q 'thrust\|blockIdx' "$TMP/ref_kernels.inc"; then
    echo "build_ref: CUDA-only code leaked into the host include" >&2; exit 1
fi
SRCS="$HERE/ref_driver.cpp $REF/FileDump.cpp $REF/FASTQData.cpp $REF/FASTQFileReader.cpp \
      $REF/InputFileHandler.cpp $REF/SortedKMerFile.cpp $REF/KMerFileMerger.cpp"
CXXFLAGS="-O2 -std=c++11 -w -fPIC -I$REF -I$TMP"
g++ $CXXFLAGS -shared -o "$OUT/libkc_ref.so" $SRCS -lpthread
g++ $CXXFLAGS -DREF_MAIN -o "$OUT/ref_count" $SRCS -lpthread
echo "build_ref: built $OUT/libkc_ref.so and $OUT/ref_count from $REF"
# The reference's GPU seam, unmodified, for sm_100a: GPUHandler.cu + FileDump.cpp + FASTQData.cpp + our driver.
if command -v nvcc >/dev/null 2>&1; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O2 -w -I"$REF" -o "$OUT/ref_gpu" \
        "$HERE/ref_gpu_driver.cpp" "$REF/GPUHandler.cu" "$REF/FileDump.cpp" "$REF/FASTQData.cpp"
    echo "build_ref: built $OUT/ref_gpu (reference GPUHandler.cu for sm_100a)"
fi
