#!/bin/bash
# ncu evidence for one round: launch list of the bench command + full captures of chosen conv launches.
# usage (under gpurun, repo root): bash tools/profile.sh <tag> [extra bench args]
TAG=${1:-r01}; shift
ARGS="--steps 2 --warmup 3 --no-cpu-baseline $@"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
# conv_tc launches per forward (in order): enc1.3 enc2.0 enc2.3 enc3.0 enc3.3 enc4.0 enc4.3 bott.0 bott.3 up4 dec4.0 dec4.3 up3
# dec3.0 dec3.3 up2 dec2.0 dec2.3 up1 dec1.0 dec1.3 = 21; skip 3 warm-up forwards (63) then take dec4.0 (idx 10), dec1.0, dec1.3
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 73 -c 1 -o gpurun_out/prof_dec40_$TAG -f \
    python bench.py $ARGS > gpurun_out/ncu_full1_$TAG.log 2>&1
echo "full dec4.0 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 82 -c 2 -o gpurun_out/prof_dec1_$TAG -f \
    python bench.py $ARGS > gpurun_out/ncu_full2_$TAG.log 2>&1
echo "full dec1.x rc=$?"
ls -la gpurun_out | tail -20
