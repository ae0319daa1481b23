#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU).  Every ncu command follows a plain run of the same command.
# The .ncu-rep files (45 MB each with --import-source) are summarised ON the box (raw page / source page as CSV) and
# removed: gpurun_out/ must stay under 64 MiB to travel back.
set -u
O=gpurun_out
WHAT=${1:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = launches ]; then
B="python bench.py --steps 2 --warmup 3 --no-cuda-graphs --no-extras --no-cpu-baseline"
$B > $O/r2_plain_bench.json 2> $O/r2_plain_bench.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $O/r2_launches.csv $B > $O/r2_ncu_bench.log 2>&1
fi
if [ "$WHAT" = all ] || [ "$WHAT" = conv ]; then
F="python tools/fused_layer_times.py 256"
$F > $O/r2_layer_times.jsonl 2> $O/r2_layer_times.err &&
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 60 -c 20 -o $O/r2_prof_conv -f $F > $O/r2_ncu_conv.log 2>&1
ncu -i $O/r2_prof_conv.ncu-rep --page raw --csv > $O/r2_conv_raw.csv 2>/dev/null
ncu -i $O/r2_prof_conv.ncu-rep --page source --csv --print-source sass -s 1 -c 1 > $O/r2_conv_src_layer1conv1.csv 2>/dev/null
ncu -i $O/r2_prof_conv.ncu-rep --page source --csv --print-source sass -s 2 -c 1 > $O/r2_conv_src_layer1conv2.csv 2>/dev/null
ncu -i $O/r2_prof_conv.ncu-rep --page source --csv --print-source sass -s 13 -c 1 > $O/r2_conv_src_layer3conv1.csv 2>/dev/null
ncu -i $O/r2_prof_conv.ncu-rep --page source --csv --print-source sass -s 0 -c 1 > $O/r2_conv_src_stem.csv 2>/dev/null
rm -f $O/r2_prof_conv.ncu-rep
fi
if [ "$WHAT" = all ] || [ "$WHAT" = i8 ]; then
I="env TQ_PROBE_TIMEOUT=900 python tools/forward_probe.py 256 i8"
$I > $O/r2_i8_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:conv_igemm -s 21 -c 19 -o $O/r2_prof_i8 -f $I > $O/r2_ncu_i8.log 2>&1
ncu -i $O/r2_prof_i8.ncu-rep --page raw --csv > $O/r2_i8_raw.csv 2>/dev/null
rm -f $O/r2_prof_i8.ncu-rep
fi
if [ "$WHAT" = all ] || [ "$WHAT" = dw ]; then
M="python tools/mobilenet_bench.py 512"
$M > $O/r2_mobilenet_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:depthwise -s 51 -c 17 -o $O/r2_prof_dw -f $M > $O/r2_ncu_dw.log 2>&1
ncu -i $O/r2_prof_dw.ncu-rep --page raw --csv > $O/r2_dw_raw.csv 2>/dev/null
ncu -i $O/r2_prof_dw.ncu-rep --page source --csv --print-source sass -s 2 -c 1 > $O/r2_dw_src.csv 2>/dev/null
rm -f $O/r2_prof_dw.ncu-rep
fi
ls -la $O | head -40; du -sh $O
