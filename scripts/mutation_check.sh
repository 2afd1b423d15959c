#!/bin/bash
# Proves that the bf16 score-parity suite can fail: builds the library with the ref/alt lanes of the fused
# score epilogue swapped on purpose (-DSB_MUTANT_SWAP_LANES, sb_tc.cu) and runs the suite against it.
# Usage (authoring container): bash scripts/mutation_check.sh build ; (GPU box): bash scripts/mutation_check.sh run
set -u
cd "$(dirname "$0")/.."
MUT=selene_b200/libselene_b200_mutant.so
if [ "${1:-build}" == "build" ]; then
  make -C selene_b200/csrc OUT=../libselene_b200_mutant.so FLAGS_EXTRA=-DSB_MUTANT_SWAP_LANES > /dev/null 2>&1 && echo "built $MUT"
else
  mkdir -p gpurun_out
  SELENE_B200_LIB=$PWD/$MUT timeout 900 python -m pytest tests/test_gpu_score_parity.py -q -m gpu -p no:cacheprovider \
      > gpurun_out/mutation_check.log 2>&1
  rc=$?
  echo "mutant suite exit code $rc (must be non-zero)" | tee -a gpurun_out/mutation_check.log
  grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/mutation_check.log | tail -8
  [ $rc -ne 0 ]
fi
