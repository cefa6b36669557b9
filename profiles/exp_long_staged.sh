#!/bin/bash
# long-RIR kernel: partitions staged through the bulk-copy engine (default) vs plain loads (-DSS_LONG_STAGED=0)
(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -2)
echo "staged:"; timeout 200 python profiles/time_long_rir.py 2>&1 | tail -3
SS_EXTRA_NVCC="-DSS_LONG_STAGED=0" python -c "from sonicsim_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -1
echo "plain loads:"; timeout 200 python profiles/time_long_rir.py 2>&1 | tail -3
python -c "from sonicsim_b200 import _lib; _lib.build(force=True)"
