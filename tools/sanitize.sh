#!/bin/bash
# compute-sanitizer over a representative subset of the GPU tests (memcheck: global/shared out-of-bounds;
# racecheck: shared-memory hazards between the warps of a CTA; synccheck: barrier misuse).
# usage (on a GPU box): bash tools/sanitize.sh [memcheck|racecheck|synccheck ...]
set -u
SEL='test_stream_cuda_vs_oracle or test_prims_cuda_vs_oracle or test_fragment_path_dense_overlap or test_stream_phong or (test_product_matches_golden and (c1-gears-f0 or c5-batch or c3-phong-arrays or prims-thick or micro-blend1 or api-gouraud-backmat or api-everything or api-fog-exp2-opaque or api-pixel-layouts-viewport or c4-overdraw-alpha-depth-two-state or c4-overdraw-add-screen-tinted-clamp-rgb8 or c2-textured-arrays-rewritten-f2 or c2-textured-bilinear-clamp-rgb8 or micro-target-bgra-fbo or texfmt-red-ubyte-nearest or texfmt-luma-half-bilinear or texfmt-bgra-float-bilinear or conform-blend-depth or examples-arrays-all-types-f0)) or test_surface_operations_stay_on_the_device'
for tool in "${@:-memcheck racecheck}"; do
  for t in $tool; do
    echo "== compute-sanitizer --tool $t"
    compute-sanitizer --tool $t --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error:|hazard|Invalid|at .*cuh?:" | head -30
  done
done
