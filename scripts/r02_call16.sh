#!/bin/bash
mkdir -p gpurun_out
python scripts/danq_config4.py 10000 > gpurun_out/danq_config4.json 2> gpurun_out/danq_config4.err; echo "danq exit $?"; cat gpurun_out/danq_config4.json; tail -2 gpurun_out/danq_config4.err
python bench_configs.py --only sweep,ism > gpurun_out/configs_sweep.jsonl 2> gpurun_out/configs_sweep.err; echo "configs exit $?"; cut -c1-330 gpurun_out/configs_sweep.jsonl
