#!/bin/bash
# Builds librtb200 variants that differ in compile-time knobs of rtb_wavefront.cu (the other objects
# are shared with the default build) into ray_tracing-rendering_b200/variants/ (git-ignored, travels
# to the GPU box).  tools/variant_sweep.py times them through RTB200_LIBRARY.
#   tools/build_variants.sh name1 "-DKNOB=1 ..." [name2 "..."] ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/ray_tracing-rendering_b200/csrc
OUT=$ROOT/ray_tracing-rendering_b200/variants
mkdir -p "$OUT"
make -C "$CSRC" -j4 >/dev/null
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    (
    nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$CSRC" \
        -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v --use_fast_math $flags \
        -c "$CSRC/rtb_wavefront.cu" -o "$OUT/wf_$name.o" 2> "$OUT/wf_$name.ptxas.log"
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/librtb200_$name.so" "$CSRC/build/rtb_api.o" \
        "$OUT/wf_$name.o" "$CSRC/build/rtb_batch_f32.o" "$CSRC/build/rtb_batch_f64.o" "$CSRC/build/rtb_multi.o" -lcudart_static -ldl \
        -Xlinker --exclude-libs=ALL
    rm -f "$OUT/wf_$name.o"
    echo "built $name ($flags)"
    ) &
done
wait
