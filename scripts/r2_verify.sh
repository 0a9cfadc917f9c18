#!/bin/bash
# First GPU run of the r2-prep prototypes (none of this CUDA has executed yet).  Run under gpurun from the branch's tree:
#   gpurun --timeout 900 -- 'bash scripts/r2_verify.sh'
# Order = cheapest and most independent first; each block prints its own verdict so one failure does not hide the rest.
set -u
mkdir -p gpurun_out
run() { echo "=== $1"; shift; timeout 300 "$@" 2>&1 | tail -8; }
run "whole GPU suite of main's tests (radix sort with fused offsets, sparse slab tables are in here)" python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_narrow.py
run "geometry kernel vs host fixture"              python -m pytest tests/test_gpu_narrow.py -q -k "device_manifolds or without_aabbs"
run "solver from edge-indexed storage vs CSR"      python -m pytest tests/test_gpu_narrow.py -q -k "edge_indexed"
run "device-resident world vs ordinary GPU world"  python -m pytest tests/test_gpu_narrow.py -q -k "device_resident_world"
run "broad phase timing (fused radix offsets)"     python bench.py --steps 20 --no-cpu
