#!/bin/bash
# Build libjt_vm.so for sm_100a (B200) in-tree. Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRCS="lib.cu march.cu vm_gather.cu vm_scatter.cu composite.cu shade.cu blur.cu pose_rays.cu field_sweep.cu field_maint.cu image_prep.cu"
for f in shade_tc.cu shade_tc_bwd.cu app_basis_tc.cu head_mlp_tc.cu; do [ -f $f ] && SRCS="$SRCS $f"; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared -Xptxas -v "$@" -o libjt_vm.so $SRCS
