#!/bin/bash
# compute-sanitizer over one small invocation of every kernel family (racecheck / memcheck).
# The round kernels use last-block-done tickets and host-mapped result flags (csrc/common.cuh); racecheck covers the
# shared-memory stages, memcheck every global access.  Usage (B200 box):  bash tools/sanitize.sh [outdir]
# Logs: <outdir>/sanitize_{memcheck,racecheck}_<family>.log ; summary lines are collected into <outdir>/sanitize_summary.txt
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SAN=/usr/local/cuda/bin/compute-sanitizer
export GKR_SANITIZE=1
declare -A FAM
FAM[smoke]="__graft_entry__.py smoke"
FAM[dense]="-m pytest -q -x tests/test_gpu_dense_sumcheck.py -k 'rounds or proof_bytes or extreme or kernel_switch' -m gpu"
FAM[deg2]="-m pytest -q -x tests/test_gpu_deg2.py -m gpu"
FAM[maps]="-m pytest -q -x tests/test_gpu_maps.py -m gpu"
FAM[msm]="-m pytest -q -x tests/test_gpu_msm.py -k 'random or edge or projective or skewed or signed_digit' -m gpu"
# pre-launched rounds (mailbox kernels, common.cuh): whole sumchecks through gkr_sumcheck_prove
FAM[mailbox]="-m pytest -q -x tests/test_gpu_dense_sumcheck.py tests/test_gpu_deg2.py -k 'proof_bytes or prover_verifier' -m gpu"
FAM[commit]="-m pytest -q -x tests/test_gpu_commit_ops.py -m gpu"
FAM[pippenger]="-m pytest -q -x tests/test_gpu_pippenger.py -k 'device_vs_oracle and 2-3-6-0' -m gpu"
: > "$OUT/sanitize_summary.txt"
for tool in memcheck racecheck; do
  for fam in smoke dense deg2 maps msm commit pippenger mailbox; do
    log="$OUT/sanitize_${tool}_${fam}.log"
    eval timeout 900 $SAN --tool $tool --print-limit 20 python ${FAM[$fam]} > "$log" 2>&1
    rc=$?
    echo "$tool $fam rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1) : $(grep -E 'passed|failed|smoke ok' "$log" | tail -1)" | tee -a "$OUT/sanitize_summary.txt"
  done
done
