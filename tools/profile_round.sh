# dev helper (GPU box): the round's ncu evidence -- launch list of one bench command, then `--set full` of one step's kernels.
# usage: bash tools/profile_round.sh <tag>     (writes gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.ncu-rep)
tag=$1
K='regex:cascade_umma|cqt_|l0_semitone|pc8_umma|upsixth|p2p|semi_umma|semitone_pool|pc2pc_umma|equiv_umma|head_|decode_kernel|peak_'
cmd="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
tail -1 gpurun_out/${tag}_ncu1.log | cut -c1-160
# one step = 23 launches of this library; bench runs 3 warm-up steps first: the 4th step is the timed one
cmd1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$cmd1 > gpurun_out/${tag}_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -s 69 -c 23 -o gpurun_out/${tag}_full $cmd1 > gpurun_out/${tag}_ncu2.log 2>&1
tail -2 gpurun_out/${tag}_ncu2.log | cut -c1-160
