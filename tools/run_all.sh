# dev helper (GPU box): GPU test suite + device-resident bench summary.  usage: bash tools/run_all.sh <log-name>
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/$1.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/$1.log").read().strip().splitlines()[-1])
c=d["stages"]["cqt"]
print(d["value"], d["ms_per_step"], "cqt", c["ms_per_step"], c["decimate_ms"], c["bank_ms"], "pcn", d["stages"]["pcn"]["ms_per_step"], d["stages"]["pcn"]["sections_ms"])
PY
