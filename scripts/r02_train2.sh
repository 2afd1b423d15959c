#!/bin/bash
# training tests, A/B of the training step by environment (same box), then the launch list of three steps
bash scripts/r02_train.sh "$@"
bash scripts/r02_train_prof.sh 2>&1 | tail -32
