#!/bin/bash
# compute-sanitizer passes over a reduced -m gpu subset that reaches every kernel template
# (tile / tile5 / window / CSR aggregation, tcgen05 projections PM_SPLIT / PLAIN / MLP0 / HEADS incl. the
# 2-CTA multicast pair, bf16 projection, graph build, stitch, vr_post).  Logs -> gpurun_out/sanitize_<tool>.log
# usage (on the GPU box): tools/sanitize.sh [tag] [tools...]
TAG=${1:-r2}
shift
TOOLS=${@:-"memcheck synccheck racecheck initcheck"}
mkdir -p gpurun_out
SUBSET='test_forward_vs_reference_module_golden or test_ragged_batch_forward_equals_per_grid_forward or test_forward_foreign_data_coo_path
 or (test_umma_matches_fp64 and (256-64-256-4 or 513-64-64-1 or 300-32-32-1)) or (test_umma_presplit_is_bit_identical and (256-256-256-4 or 513-64-64-1))
 or (test_proj_bf16_tcgen05_vs_fp64 and 256-64-4-77) or (test_forward_bf16_default_shapes and 8-connected-64-4)
 or (test_radius2_tiled_aggregation_vs_csr_kernel and 64-4-24)
 or test_tile_merger_bit_exact_vs_reference or test_pipeline_process_vs_reference or test_native_vr_batching_vs_reference
 or test_sidecar_batch_bit_exact_vs_reference_golden or test_vr_apply_results_bit_exact_vs_oracle
 or (test_graph_vs_reference_golden and (g64_holes or g33_k24 or g48_4conn)) or test_ragged_batch_equals_per_grid'
SUBSET=$(echo $SUBSET)
EXTRA_SUBSET=${SANITIZE_EXTRA:-}
[ -n "$EXTRA_SUBSET" ] && SUBSET="$SUBSET or $EXTRA_SUBSET"
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${TAG}_${tool}.log
  opts="--tool $tool --target-processes all --print-limit 30 --error-exitcode 66"
  [ "$tool" = memcheck ] && opts="$opts --leak-check no"
  [ "$tool" = initcheck ] && opts="$opts --track-unused-memory no"
  echo "=== $tool ===" | tee $log
  start=$(date +%s)
  timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer $opts \
      python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$SUBSET" > gpurun_out/sanitize_${TAG}_${tool}.full 2>&1
  rc=$?
  end=$(date +%s)
  # keep the summary small: pytest tail, sanitizer summary lines, first errors
  { echo "exit code $rc, $((end-start)) s";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_${TAG}_${tool}.full | tail -12;
    grep -E -m 40 "=========( |     )(Invalid|Uninitialized|Race|Barrier|Error|Warning|Hazard|at |by )" gpurun_out/sanitize_${TAG}_${tool}.full; } >> $log
  head -c 300000 gpurun_out/sanitize_${TAG}_${tool}.full > gpurun_out/sanitize_${TAG}_${tool}.head; rm -f gpurun_out/sanitize_${TAG}_${tool}.full
  tail -8 $log
done
