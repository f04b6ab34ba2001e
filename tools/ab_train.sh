# dev helper (GPU box): A/B of two library builds on the training step.  usage: bash tools/ab_train.sh <variant.so name under tools/bin>
for rep in 1 2; do for v in default $1; do
  if [ "$v" = "default" ]; then unset AKE_LIB_PATH; else export AKE_LIB_PATH=$PWD/tools/bin/$v.so; fi
  for b in 8 64; do timeout 200 python bench.py --train --batch $b --steps 100 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v B=$b', round(d['ms_per_step'],3), 'ms')"; done
done; done
