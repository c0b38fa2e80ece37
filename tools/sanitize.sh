#!/bin/bash
# compute-sanitizer memcheck + racecheck over short rollouts of the product path.
# Usage (GPU box): bash tools/sanitize.sh <out_dir>
out=${1:-gpurun_out}
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
run() { # tool tag env_id batch steps
  timeout 900 $CS --tool "$1" --print-limit 400 --error-exitcode 9 \
    python tools/prof_rollout.py "$3" "$4" "$5" > "$out/sanitizer_${1}_${2}.log" 2>&1
  echo "$1 $2 rc=$?" >> "$out/sanitizer_summary.txt"
  tail -3 "$out/sanitizer_${1}_${2}.log" >> "$out/sanitizer_summary.txt"
}
: > "$out/sanitizer_summary.txt"
run memcheck cluster256 ClusterColour-Demo-LoRes4E-v0 256 30
run memcheck mrpool64 MatchRegions-TestAll-LoResStack-v0 64 30
run racecheck cluster256 ClusterColour-Demo-LoRes4E-v0 256 12
run racecheck mrpool64 MatchRegions-TestAll-LoResStack-v0 64 12
cat "$out/sanitizer_summary.txt"
