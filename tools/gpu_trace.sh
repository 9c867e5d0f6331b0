#!/bin/bash
# Builds the pointwise kernel with the timeline trace compiled in, runs tools/trace_pw.py, restores nothing
# (the box is discarded); run only through gpurun.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
L=x3d_tf_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr -DX3D_PW_TRACE -c x3d_tf_b200/csrc/x3d_pw_tc.cu -o $L/x3d_pw_tc.o 2>&1 | grep -v deprecated
nvcc -shared -o $L/libx3d_b200.so $L/*.o -cudart static 2>&1 | grep -v deprecated
python tools/trace_pw.py
