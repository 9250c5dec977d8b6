#!/bin/bash
# GPU-box recipe behind the files in profiles/ (run through gpurun from the repo root):
#   gpurun --timeout 1700 -- 'bash profiles/run_profile.sh r01'
# 1. bench line (never under a profiler) with nvidia-smi clocks beside it, 2. launch list of the same command,
# 3. one --set full capture of the scan kernel (and of the select / LUT kernels).
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $out/clocks_$tag.csv &
SMI=$!
python bench.py --steps 50 --warmup 5 > $out/bench_$tag.json 2> $out/bench_$tag.err
kill $SMI
tail -c 4000 $out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|DeviceRadix|DeviceScan' -s 40 -c 120 --csv \
    --log-file $out/launches_$tag.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan_pk -s 4 -c 2 -f -o $out/scan_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_select|k_lut_reg|k_lut_quant' -s 8 -c 3 -f -o $out/aux_$tag \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_aux_$tag.log 2>&1
ls -la $out
