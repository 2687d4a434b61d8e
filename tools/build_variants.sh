#!/bin/bash
# Builds kernel launch-shape variants of the product library side by side (round-2 sweep):
#   tools/build_variants.sh tag "KDEFS" [tag "KDEFS" ...]
# -> lulesh_b200/lib/variants/lib_<tag>.so (git-ignored, travels to the GPU box); select one
# with LULESH_B200_LIB=<path> (lulesh_b200/__init__.py).
set -e
cd "$(dirname "$0")/.."
mkdir -p lulesh_b200/lib/variants
while [ $# -ge 2 ]; do
   tag=$1; defs=$2; shift 2
   make -s -j8 OBJ=build_$tag LIB=lulesh_b200/lib/variants/lib_$tag.so KDEFS="$defs" lulesh_b200/lib/variants/lib_$tag.so
   grep -A2 -E "k_force|k_kinematics" build_$tag/kernels.ptxas.log | grep -E "Used|spill" | sed "s/^/[$tag] /"
done
