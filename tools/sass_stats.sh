#!/bin/bash
# Per-kernel SASS statistics of a built library: instruction count, LDL/STL count.   tools/sass_stats.sh [lib.so]
LIB=${1:-$(dirname "$0")/../cuda-csg-tree-raycasting_b200/libcsg_b200.so}
cuobjdump -sass "$LIB" | awk '/Function :/ {name=$3} /^ +\/\*[0-9a-f]+\*\/ / {n[name]++; if ($0 ~ /LDL|STL/) l[name]++} END {for (k in n) printf "%6d instr %3d LDL/STL  %s\n", n[k], l[k]+0, k}' | sort -k5
