# dev helper (GPU box, N GPUs): the multi-GPU records of a round -- headline bench, configs[3] sweep, configs[4] training step.
# usage: bash tools/multi_gpu_record.sh <N> <tag>
N=$1; tag=$2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29541 --steps 20 --warmup 3 > gpurun_out/${tag}_bench$N.log 2> gpurun_out/${tag}_bench$N.err
run 29542 --sweep --steps 3 > gpurun_out/${tag}_sweep$N.log 2> gpurun_out/${tag}_sweep$N.err
run 29543 --train --steps 30 --warmup 5 > gpurun_out/${tag}_train$N.log 2> gpurun_out/${tag}_train$N.err
python - <<PY
import json
for kind in ("bench", "sweep", "train"):
    try:
        for ln in open("gpurun_out/${tag}_%s$N.log" % kind).read().strip().splitlines():
            if not ln.startswith("{"): continue
            d = json.loads(ln)
            extra = ""
            if d.get("e2e"): extra = " e2e %.0f e2e_i16 %.0f" % (d["e2e"]["value"], (d.get("e2e_i16") or {}).get("value", 0))
            print(kind, d["n_gpus"], "gpus:", "%.0f %s" % (d["value"], d["unit"]), "%.3f ms/step" % d["ms_per_step"], d["config"].get("global_batch", ""), extra)
    except Exception as e:
        print(kind, "failed:", e)
        print(open("gpurun_out/${tag}_%s$N.err" % kind).read()[-1500:])
PY
