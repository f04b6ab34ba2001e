# dev helper (GPU box): bench summary for several library build variants (tools/bin/<name>.so).  usage: bash tools/lib_sweep.sh <tag> name1 name2 ...
tag=$1; shift
for v in "$@"; do
  if [ "$v" = "default" ]; then unset AKE_LIB_PATH; else export AKE_LIB_PATH=$PWD/tools/bin/$v.so; fi
  timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 40 > gpurun_out/${tag}_$v.log 2> gpurun_out/${tag}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$v.log").read().strip().splitlines()[-1])
    c=d["stages"]["cqt"]
    print("$v: clips/s %.0f step %.3f ms | cqt %.3f (decimate %.3f bank %.3f) | pcn %.3f" % (d["value"], d["ms_per_step"], c["ms_per_step"], c["decimate_ms"], c["bank_ms"], d["stages"]["pcn"]["ms_per_step"]), {k: round(x,3) for k,x in d["stages"]["pcn"]["sections_ms"].items()})
except Exception as e:
    print("$v failed:", e); print(open("gpurun_out/${tag}_$v.err").read()[-800:])
PY
done
