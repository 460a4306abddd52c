#!/bin/bash
# SASS of one kernel of libcsg_b200.so: tools/sass_fn.sh <substring of the mangled name>
cuobjdump -sass "$(dirname "$0")/../cuda-csg-tree-raycasting_b200/libcsg_b200.so" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on {print}'
