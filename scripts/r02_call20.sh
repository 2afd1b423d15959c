#!/bin/bash
# host-API GPU tests (writers / VEP files) in both writer modes + the file-to-file timing at the new default block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_analyze.py tests/test_vcf_ingest.py tests/test_host.py tests/test_gpu_score_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_api.log 2>&1; echo "api tests exit $?"; tail -2 gpurun_out/t_api.log
SB_WRITER_MMAP=1 timeout 600 python -m pytest tests/test_gpu_analyze.py tests/test_vcf_ingest.py -x -q -m gpu -p no:cacheprovider > gpurun_out/t_api_mmap.log 2>&1; echo "api tests (mmap writer) exit $?"; tail -2 gpurun_out/t_api_mmap.log
python - <<'PY' 2>&1 | grep -v Warning | cut -c1-500
import subprocess, sys
print(subprocess.run([sys.executable, "scripts/time_vep_file.py", "100000", "9472"], capture_output=True, text=True).stdout)
PY
