#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/score_parity_stats.txt
timeout 900 python -m pytest tests/test_gpu_score_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/t_score.log 2>&1; echo "score-parity exit $?"
tail -3 gpurun_out/t_score.log; cat gpurun_out/score_parity_stats.txt
bash scripts/mutation_check.sh run; echo "mutation check exit $?"
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref arm exit $?"; cat gpurun_out/bench_ref.json
python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
