#!/bin/bash
# compute-sanitizer over the hot path (memcheck + racecheck on every kernel).
# Usage (on the GPU box): bash profiles/sanitize.sh <out-dir>
# 1. __graft_entry__.smoke(): single-sequence calls of every stage + one batched extraction
# 2. profiles/sanitize_probe.py: one batched replay (3 sequences x 6 VLP-16 scans: association with
#    ticketed moment merges, evaluation, commit, removals) and the streaming stage-3 kernels
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  for what in smoke probe; do
    if [ $what = smoke ]; then cmd="python __graft_entry__.py smoke"; else cmd="python profiles/sanitize_probe.py"; fi
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 7 $cmd > "$OUT/sanitize_${tool}_${what}.log" 2>&1
    echo "$tool $what exit=$?" >> "$OUT/sanitize_summary.txt"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|probe ok" "$OUT/sanitize_${tool}_${what}.log" >> "$OUT/sanitize_summary.txt"
  done
done
cat "$OUT/sanitize_summary.txt"
