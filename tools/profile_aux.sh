#!/bin/bash
# ncu evidence for the non-tensor stages (rolling ball, labelling + table, overlay): per-launch metrics of one warm
# repetition plus one `--set full` capture (with source) of the two dominant kernels.
# usage (under gpurun, repo root): bash tools/profile_aux.sh <tag>
TAG=${1:-r02}
M="gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__shared_mem_per_block_static,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.avg"
mkdir -p gpurun_out
python tools/aux_driver.py --reps 3 --overlay > gpurun_out/aux_plain_$TAG.txt 2>&1 || { echo "plain run failed"; tail gpurun_out/aux_plain_$TAG.txt; exit 1; }
cat gpurun_out/aux_plain_$TAG.txt
# launches of the third (warm) repetition: 3 rolling ball + 7 label + 5 overlay = 15 per repetition
ncu --metrics $M --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/aux_metrics_$TAG.csv \
    python tools/aux_driver.py --reps 3 --overlay > /dev/null 2>&1; echo "aux metrics rc=$?"
ncu --set full --import-source on --clock-control none -k regex:'morph_chord' -s 4 -c 2 -o gpurun_out/morph_$TAG -f \
    python tools/aux_driver.py --reps 3 > /dev/null 2>&1; echo "morph full rc=$?"
ncu --set full --import-source on --clock-control none -k regex:'ccl_tile' -s 2 -c 1 -o gpurun_out/ccl_tile_$TAG -f \
    python tools/aux_driver.py --reps 3 > /dev/null 2>&1; echo "ccl full rc=$?"
du -sh gpurun_out
