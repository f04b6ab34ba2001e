# dev helper (GPU box): CQT stage time as a function of the clips per L2-resident group.  usage: bash tools/cqt_group_sweep.sh <tag> g1 g2 ...
tag=$1; shift
for g in "$@"; do
  if [ "$g" = "auto" ]; then unset AKE_CQT_GROUP; else export AKE_CQT_GROUP=$g; fi
  timeout 200 python bench.py --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/${tag}_g$g.log 2> gpurun_out/${tag}_g$g.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_g$g.log").read().strip().splitlines()[-1])
    c=d["stages"]["cqt"]
    print("group $g: clips/s %.0f step %.3f ms | cqt %.3f (decimate %.3f bank %.3f) | pcn %.3f launches/step %d" % (d["value"], d["ms_per_step"], c["ms_per_step"], c["decimate_ms"], c["bank_ms"], d["stages"]["pcn"]["ms_per_step"], d["gpu_launches"]//d["steps"]))
except Exception as e:
    print("group $g failed:", e); print(open("gpurun_out/${tag}_g$g.err").read()[-800:])
PY
done
