#!/bin/bash
# compute-sanitizer evidence for the synchronisation protocols of the sweep kernels (run on the GPU box):
#   racecheck (shared-memory hazards) and memcheck on small instances of
#   - the chain kernels (wavefront flags in shared memory): k_sweep, k_sweep<RS>, k_sweep_cb
#   - the cluster kernels (DSMEM slots, mbarriers, st.release.cluster flags): k_sweep_blk, k_sweep_slice_cl
#   - the CTA-per-slice kernels (global st.release / ld.acquire flags + ticket): k_sweep_slice_ws, k_sweep_cc, _cc3
#   - the dataflow case-control kernel (global-memory records with sentinels, atomic tickets): k_sweep_ccd
# Output: gpurun_out/sanitize_<tool>.log (copied to profiles/ by hand).
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -n 5 gpurun_out/sanitize_$tool.log
done
