#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference (deepseek-ai/DeepGEMM) into oracle/_ref/
# so that GPU tests / the A-B bench can run the reference's own SM100 kernel beside ours.
#
# Nothing from /root/reference is copied into tracked paths: oracle/_ref/ is git-ignored (it is an install
# directory, exactly what `pip install --target` of the reference would produce: its python package, the
# JIT include tree the reference needs at run time, and the compiled host extension `_C`).
# It is NOT gpurun-ignored, so it travels to the GPU box like our own built .so files.
#
# Recipe = SURVEY.md Appendix D (the reference's own setup.py fails in this image without
# `-include cuda_fp8.h`), so we compile its single host source directly.
set -euo pipefail
REF_SRC=${REF_SRC:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF_SRC" ]; then
    echo "[build_ref] $REF_SRC not present (GPU box?) -- using prebuilt oracle/_ref if any"; exit 0
fi
if [ -f "$OUT/deep_gemm/_C.cpython-312-x86_64-linux-gnu.so" ] && [ -z "${FORCE:-}" ]; then
    echo "[build_ref] already built"; exit 0
fi
rm -rf "$OUT"; mkdir -p "$OUT"
# 1. the install image (python package + JIT headers), as pip --target would lay it out
cp -r "$REF_SRC/deep_gemm" "$OUT/deep_gemm"
rm -rf "$OUT/deep_gemm/include/cute" "$OUT/deep_gemm/include/cutlass"
cp -r "$REF_SRC/third-party/cutlass/include/cute" "$OUT/deep_gemm/include/cute"
cp -r "$REF_SRC/third-party/cutlass/include/cutlass" "$OUT/deep_gemm/include/cutlass"
# setup.py generates deep_gemm/envs.py with persisted env defaults; an empty one is equivalent
[ -f "$OUT/deep_gemm/envs.py" ] || echo "persistent_envs = dict()" > "$OUT/deep_gemm/envs.py"
# 2. the host extension, compiled from the source where it lies
T=$(python -c "import torch,os;print(os.path.dirname(torch.__file__))")
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
CUDA=${CUDA_HOME:-/usr/local/cuda}
TMP=$(mktemp -d)
( cd "$REF_SRC" && g++ -std=c++17 -O2 -fPIC -include cuda_fp8.h -Wno-psabi -Wno-deprecated-declarations \
    -D_GLIBCXX_USE_CXX11_ABI=1 -DTORCH_EXTENSION_NAME=_C -DTORCH_API_INCLUDE_EXTENSION_H \
    -I$CUDA/include -Ideep_gemm/include -Ithird-party/cutlass/include -Ithird-party/fmt/include \
    -I$T/include -I$T/include/torch/csrc/api/include -I$PYINC \
    -c csrc/python_api.cpp -o $TMP/python_api.o )
g++ -shared $TMP/python_api.o -o "$OUT/deep_gemm/_C.cpython-312-x86_64-linux-gnu.so" \
    -L$T/lib -L$CUDA/lib64 -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
    -lcudart -lnvrtc -lcublasLt -Wl,-rpath,$T/lib
rm -rf $TMP
echo "[build_ref] built $OUT"
