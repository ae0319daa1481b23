#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU).  Every ncu command follows a plain run of the same command.
# Outputs land in gpurun_out/; summaries are extracted on the CPU box with tools/ncu_summary.py.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cuda-graphs --no-extras --no-cpu-baseline"
$B > $O/r2_plain_bench.json 2> $O/r2_plain_bench.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r2_launches.csv $B > $O/r2_ncu_bench.log 2>&1
F="python tools/fused_layer_times.py 256"
$F > $O/r2_layer_times.jsonl 2> $O/r2_layer_times.err &&
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 20 -o $O/r2_prof_conv -f $F > $O/r2_ncu_conv.log 2>&1
I="env TQ_PROBE_TIMEOUT=900 python tools/hang_probe.py 256 i8"
$I > $O/r2_i8_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 21 -c 19 -o $O/r2_prof_i8 -f $I > $O/r2_ncu_i8.log 2>&1
M="python tools/mobilenet_bench.py 512"
$M > $O/r2_mobilenet_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:depthwise -s 51 -c 17 -o $O/r2_prof_dw -f $M > $O/r2_ncu_dw.log 2>&1
ls -la $O/*.ncu-rep
