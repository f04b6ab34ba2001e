# dev helper (GPU box, N GPUs): configs[4] records -- training step at batch 8 and 64 per GPU.  usage: bash tools/train_record.sh <N> <tag>
N=$1; tag=$2
run() { if [ "$N" = "1" ]; then python bench.py "${@:2}"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; fi; }
for b in 8 64; do
  run $((29500 + b)) --train --batch $b --steps 100 --warmup 5 > gpurun_out/${tag}_train${N}_b$b.log 2> gpurun_out/${tag}_train${N}_b$b.err
  tail -1 gpurun_out/${tag}_train${N}_b$b.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train', d['n_gpus'], 'gpus batch $b/GPU:', round(d['value']), 'clips/s', round(d['ms_per_step'],3), 'ms/step')" || tail -5 gpurun_out/${tag}_train${N}_b$b.err
done
