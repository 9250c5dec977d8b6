#!/bin/bash
# GPU-box recipe behind the r02_* files in profiles/ (run through gpurun from the repo root, ONE GPU):
#   gpurun --timeout 2400 -- 'bash profiles/run_profile_r02.sh'
# Nothing printed under ncu / compute-sanitizer is a bench value; the bench lines come from the plain runs.
out=gpurun_out
mkdir -p $out
# 1. launch list (per-launch durations, --clock-control none) of the headline command
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|DeviceRadix|DeviceScan|DeviceSegmented' -s 60 -c 160 --csv \
    --log-file $out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustain-s 0 > $out/r02_ncu_launch.log 2>&1
# 2. --set full: the batched scan (headline), select / LUT kernels
ncu --set full --clock-control none --import-source on -k regex:k_scan_pk -s 6 -c 2 -f -o $out/r02_scan_pk \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-s 0 > $out/r02_ncu_scan_pk.log 2>&1
# 3. --set full: the low-batch (HBM-bound) scan, single query ranking 4M codes of 16 and 32 bytes
ncu --set full --clock-control none --import-source on -k regex:k_scan1 -s 6 -c 2 -f -o $out/r02_scan1 \
    python profiles/hbm_scan_probe.py 4000000 > $out/r02_ncu_scan1.log 2>&1
# 4. --set full: encode kernels (tensor-pipe % of the rotation, fp32 pipe of the fine argmin)
ncu --set full --clock-control none --import-source on -k regex:'k_rotate_dmma|k_fine_argmin32|k_coarse_assign' -s 3 -c 3 -f -o $out/r02_encode \
    python bench.py --config c5 --n-db 8000000 --steps 1 --warmup 1 --no-cpu-baseline > $out/r02_ncu_encode.log 2>&1
# 5. memcheck + racecheck of the smoke run (search + encode through the C-ABI)
compute-sanitizer --tool memcheck --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > $out/r02_sanitizer_memcheck.log 2>&1
compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > $out/r02_sanitizer_racecheck.log 2>&1
tail -5 $out/r02_sanitizer_memcheck.log $out/r02_sanitizer_racecheck.log
ls -la $out | tail -20
