#!/bin/bash
# GPU-box call: in-kernel cycle breakdown of the preconditioned solve on the listed workloads
mkdir -p gpurun_out
for w in ${1:-dambreak2d_1m dambreak3d_1m}; do
  timeout 600 python scripts/cg_probe.py $w > gpurun_out/probe_$w.log 2>&1; grep -E "per iteration|us_per_iter|ms_per_step|iters_last|precond|CTA 0|rror" gpurun_out/probe_$w.log
done
