#!/bin/bash
# Round-2 evidence run on one B200: bench lines (c4 default, c3, c5), ncu launch lists and full captures.
# usage (on the GPU box, repo root): bash scripts/gpu_round2_evidence.sh <tag>
tag=${1:-r2x}
o=gpurun_out
timeout -s KILL 400 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
timeout -s KILL 400 python bench.py --config c5 > $o/${tag}_bench_c5.json 2> $o/${tag}_bench_c5.err
timeout -s KILL 300 python bench.py --config c3 > $o/${tag}_bench_c3.json 2> $o/${tag}_bench_c3.err
# launch lists (serialised, cold cache): shares of the step
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/${tag}_c4_launch_list.csv \
    python scripts/profile_step.py --config c4 --steps 3 > $o/${tag}_c4_launch.log 2>&1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/${tag}_c5_launch_list.csv \
    python scripts/profile_step.py --config c5 --steps 2 > $o/${tag}_c5_launch.log 2>&1
# full captures: the wide MLP kernels of one c5 step, the kernels of one c4 step
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tcw -c 4 -f -o $o/${tag}_c5_tcw \
    python scripts/profile_step.py --config c5 --steps 1 > $o/${tag}_c5_ncu.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --launch-skip 6 -c 6 -f -o $o/${tag}_c4_step \
    python scripts/profile_step.py --config c4 --steps 2 > $o/${tag}_c4_ncu.log 2>&1
tail -c 300 $o/${tag}_bench_n1.json; echo; tail -c 300 $o/${tag}_bench_c5.json; echo
