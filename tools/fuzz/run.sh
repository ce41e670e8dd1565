#!/bin/bash
# tools/fuzz/run.sh — sanitizer fuzzing of the host-side code above and beside the C ABI (no GPU needed).
# Every run is bounded by `timeout`; a finding aborts with the sanitizer's report.  Usage: tools/fuzz/run.sh [seconds]
set -u
cd "$(dirname "$0")"
LIMIT=${1:-240}
CSRC=../../pathfinder_b200/csrc
SAN="-fsanitize=address,undefined -fno-sanitize-recover=all"
OUT=${TMPDIR:-/tmp}/pf_fuzz
mkdir -p "$OUT"
status=0
run() { # name, command...
    local name=$1; shift
    echo "== $name"
    ( cd "$OUT" && ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:hard_rss_limit_mb=6000 timeout "$LIMIT" "$@" ) | tail -2
    local rc=${PIPESTATUS[0]}
    if [ "$rc" = 124 ]; then echo "   (time limit reached without a finding)"; elif [ "$rc" != 0 ]; then echo "   FAILED ($rc)"; status=1; fi
}
g++ -std=c++17 -O1 -g $SAN font_asan.cpp $CSRC/font.cpp -o "$OUT/font_asan" && run font "$OUT/font_asan" "${FONT:-/root/reference/resources/fonts/Roboto-Regular.ttf}"
g++ -std=c++17 -O1 -g $SAN host_asan.cpp $CSRC/svg.cpp $CSRC/stroke.cpp $CSRC/dilate.cpp -o "$OUT/host_asan" && run svg+stroke+dilate "$OUT/host_asan"
nvcc -Wno-deprecated-gpu-targets -std=c++17 -O1 -g -Xcompiler -fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=all \
    -x cu scene_fuzz.cpp $CSRC/scene.cpp $CSRC/dilate.cpp -o "$OUT/scene_fuzz" && run scene-builder "$OUT/scene_fuzz"
nvcc -Wno-deprecated-gpu-targets -std=c++17 -O1 -g -Xcompiler -fsanitize=thread \
    -x cu scene_tsan.cpp $CSRC/scene.cpp $CSRC/dilate.cpp -o "$OUT/scene_tsan" && run scene-builder-threads "$OUT/scene_tsan"
# The scene proxy (worker thread, one-command hand-over, teardown with a build in flight) under ThreadSanitizer; the whole
# library is linked in (the proxy lives next to the renderer), no GPU is touched.
nvcc -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -std=c++17 -O1 -g -fmad=false -Xcompiler -fsanitize=thread \
    -x cu proxy_tsan.cpp $CSRC/kernels.cu $CSRC/composite.cu $CSRC/renderer.cu $CSRC/scene.cpp $CSRC/stroke.cpp $CSRC/svg.cpp \
    $CSRC/dilate.cpp $CSRC/font.cpp -o "$OUT/proxy_tsan" -ldl 2>/dev/null && run scene-proxy-threads "$OUT/proxy_tsan"
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -msse4.1 -pthread $SAN -shared -o "$OUT/libpf_oracle_asan.so" ../../oracle/pf_oracle.cpp && \
    LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" run oracle python "$PWD/oracle_asan.py" "$OUT/libpf_oracle_asan.so"
# The CPU test suite against a build of the product library whose host code is instrumented (device code untouched).
( cd $CSRC && nvcc -Wno-deprecated-gpu-targets -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-O1,-ffp-contract=off,-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=all \
    -shared -o "$OUT/libpf_cuda_asan.so" kernels.cu composite.cu renderer.cu -x cu scene.cpp stroke.cpp svg.cpp dilate.cpp font.cpp 2>/dev/null ) && \
    LIMIT=$((LIMIT > 600 ? LIMIT : 600)) PF_CUDA_LIB="$OUT/libpf_cuda_asan.so" \
    LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
    run cpu-suite-on-sanitizer-build python -m pytest "$PWD/../../tests" -x -q -m "not gpu" -p no:cacheprovider \
        --deselect "$PWD/../../tests/test_strip_gloo.py"
exit $status
