#!/bin/bash
# round 2, GPU call 1: new bf16 score-parity suite (+ mutation check), smoke, the whole -m gpu suite, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_score_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/t_score.log 2>&1; echo "score-parity exit $?"
tail -30 gpurun_out/t_score.log
bash scripts/mutation_check.sh run; echo "mutation check exit $?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_score_parity.py > gpurun_out/t_all.log 2>&1; echo "all gpu tests exit $?"
tail -5 gpurun_out/t_all.log
python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
