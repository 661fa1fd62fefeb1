#!/bin/bash
# A/B builds of the CUDA library (compile-time switches) into raytracer-ceng477-graphics-hw-1_b200/ab/<name>.so for
# tools/kernel_ab.py.  Built files are git-ignored and travel to the GPU box with the gpurun snapshot.
#   tools/build_variants.sh name:"-DFLAG=.. -DFLAG2=.." ...
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$HERE/raytracer-ceng477-graphics-hw-1_b200"
mkdir -p "$PKG/ab"
SRCS="$PKG/csrc/render_v2.cu $PKG/csrc/assemble.cu $PKG/csrc/api.cu $PKG/csrc/scene_build.cu $PKG/csrc/bvh_lbvh.cu $PKG/csrc/bvh_sah_device.cu $PKG/csrc/bvh_reinsert.cu $PKG/csrc/ref_order_device.cu $PKG/csrc/selftest.cu $PKG/csrc/ref_order.cpp $PKG/csrc/bvh_host.cpp"
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  [ "$flags" = "$v" ] && flags=""
  echo "== $name: $flags"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -ccbin g++ \
       -I"$HERE/include" -I"$PKG/csrc" $flags -Xptxas -v -shared $SRCS -o "$PKG/ab/$name.so" -lrt 2> "$PKG/ab/$name.ptxas.log" &
done
wait
for v in "$@"; do name="${v%%:*}"; ls -la "$PKG/ab/$name.so"; grep -A1 "render_kernel_v2ILi1ELb0" "$PKG/ab/$name.ptxas.log" | grep -E "registers" | head -2; done
