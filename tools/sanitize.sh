#!/bin/bash
# compute-sanitizer over the hot path: memcheck, racecheck (shared-memory hazards of the tile rasterizer, the
# binner's scans, the mesh kernel's per-warp vertex records), initcheck and synccheck on smoke() — one small frame in both
# raster modes + resolve, checked against the oracle — and on the peer-exchange kernels that spin on flags
# (tests/test_visbuffer_gpu.py::test_send_pixels_with_flow_control_on_one_gpu) and the deferred / overdraw programs.
# usage: bash tools/sanitize.sh [tag]     -> gpurun_out/<tag>_sanitizer.txt
tag=${1:-r02}
out=gpurun_out/${tag}_sanitizer.txt
: > $out
run() {   # tool, label, command...
    tool=$1; label=$2; shift 2
    echo "=== compute-sanitizer --tool $tool : $label" >> $out
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 "$@" > /tmp/san.log 2>&1
    rc=$?
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|hazard|Invalid|Uninitialized|passed|failed|smoke ok|=========     at " /tmp/san.log | head -40 >> $out
    echo "exit code $rc" >> $out
}
run memcheck  "smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run racecheck "smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run synccheck "smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run initcheck "smoke()" python -c "import __graft_entry__ as g; g.smoke()"
run memcheck  "peer send with flow control, deferred + overdraw programs, clipper" python -m pytest -q -x -m gpu tests/test_visbuffer_gpu.py::test_send_pixels_with_flow_control_on_one_gpu "tests/test_deferred_gpu.py::test_gbuffer_bit_exact_on_scenes[direct_clip]" "tests/test_debug_programs.py::test_overdraw_program_bit_exact[binned]"
run racecheck "peer send with flow control, deferred + overdraw programs, clipper" python -m pytest -q -x -m gpu tests/test_visbuffer_gpu.py::test_send_pixels_with_flow_control_on_one_gpu "tests/test_deferred_gpu.py::test_gbuffer_bit_exact_on_scenes[direct_clip]" "tests/test_debug_programs.py::test_overdraw_program_bit_exact[binned]"
run memcheck  "scissor rows (band cull, band resolve, band GetPixels) + alpha raster" python -m pytest -q -x -m gpu tests/test_scissor_gpu.py -k "alpha and direct_clip or knot and binned or bands_fill or band_cull"
run racecheck "alpha raster (shared-memory setup / survivor list per warp)" python -m pytest -q -x -m gpu tests/test_alpha_gpu.py -k binned
cat $out
