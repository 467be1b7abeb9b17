#!/bin/bash
# compute-sanitizer over the smoke invocation of the hot path (2 frames x 2 detections, 1 + 1 iterations, RANSAC scoring
# of a 3-view scene): memcheck, racecheck (shared-memory hazards) and synccheck.  Logs go to gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  COSYB200_GRAPH=0 timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 20 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
