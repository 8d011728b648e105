#!/bin/bash
# ncu evidence for one round, small enough to come back through gpurun_out/ (< 64 MiB): the launch list of the bench
# command plus per-launch metric CSVs for every launch of one warm step.  Summarise with tools/summarize_profiles.py
# and tools/metrics_table.py (see profiles/README.md).
# usage (under gpurun, repo root): bash tools/profile.sh <tag>
TAG=${1:-r01}
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --frames 0"
M="gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg,smsp__inst_executed.sum"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail gpurun_out/plain_$TAG.err; exit 1; }
# (only this library's kernels: building the model launches ~900 small torch kernels -- the weight composition -- first)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_|stem_|morph_|stretch_|ccl_' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > /dev/null 2>&1; echo "launch list rc=$?"
# one warm forward = 18 launches (stem + 17 conv, the upconvs ride in dec{l}.0); 3 warm-up steps precede it
ncu --metrics $M --clock-control none -k regex:'conv_|stem_' -s 54 -c 18 --csv --log-file gpurun_out/fwd_metrics_$TAG.csv \
    python bench.py $ARGS > /dev/null 2>&1; echo "forward metrics rc=$?"
ncu --metrics $M --clock-control none -k regex:'morph_|stretch_|ccl_' -s 30 -c 10 --csv --log-file gpurun_out/aux_metrics_$TAG.csv \
    python bench.py $ARGS > /dev/null 2>&1; echo "aux metrics rc=$?"
du -sh gpurun_out
