timeout 300 python -m pytest tests/test_gpu_pcn.py -m gpu -x -q 2>&1 | tail -3; timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/$1.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/$1.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages"]["cqt"], d["stages"]["pcn"]["ms_per_step"], d["stages"]["pcn"]["sections_ms"])
PY
