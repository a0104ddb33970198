#!/usr/bin/env bash
# full GPU test-suite, no early stop, printed diagnostics kept (-rA)
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 2400 python -m pytest tests -m gpu -q -rA --timeout=900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|rc=" gpurun_out/${TAG}_pytest.log | tail -5
