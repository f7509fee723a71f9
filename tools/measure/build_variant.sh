#!/bin/bash
# Kernel experiments: build a variant of the library with extra nvcc flags into variants/<name>/ (git-ignored, travels
# to the GPU box); run it with DNMF_B200_LIB=variants/<name>/libdnmf_b200.so.   tools/measure/build_variant.sh base "-DDNMF_AFFINE_BODIES=0"
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p variants/$name
DNMF_B200_OUT_DIR=$PWD/variants/$name DNMF_NVCC_FLAGS="$*" python -m dnmf_b200.build --force >/dev/null
rm -rf variants/$name/obj
ls -la variants/$name/libdnmf_b200.so
