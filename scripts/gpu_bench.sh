#!/bin/bash
# One GPU call: parity tests, the bench line, the ncu launch list and one full capture of the
# tcgen05 layer kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_check.sh > gpurun_out/check.log 2>&1
grep -E "exit|passed|failed" gpurun_out/check.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" == "ncu" ]; then
  SMALL="python bench.py --steps 2 --warmup 1 --no-cpu --variants-per-step 4096 --genome-mbp 24"
  $SMALL > gpurun_out/ncu_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  $SMALL > gpurun_out/ncu_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:tc_layer -s 4 -c 4 -o gpurun_out/prof_tc $SMALL > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
