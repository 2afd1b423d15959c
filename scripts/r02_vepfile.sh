#!/bin/bash
# file-to-file VEP: block size x writer mode, one box
mkdir -p gpurun_out
python -c "import os; print('cores', os.cpu_count())"; df -h /tmp | tail -1
( SB_WRITER_MMAP=0 python scripts/time_vep_file.py 100000 32768 9472 18944
  SB_WRITER_MMAP=1 python scripts/time_vep_file.py 100000 32768 9472 18944 ) 2>&1 | grep -v Warning | tee gpurun_out/vepfile.log | cut -c1-600
