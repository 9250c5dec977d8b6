#!/bin/bash
# developer capture: --set full of the low-batch scan at 10 M rows, 16- and 32-byte codes, one query over the whole index
out=gpurun_out; mkdir -p $out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_scan1<.int.16>' -s 10 -c 1 -f -o $out/dev_scan1_16 \
    python profiles/hbm_scan_probe.py 10000000 > $out/dev_ncu_scan1_16.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_scan1<.int.32>' -s 10 -c 1 -f -o $out/dev_scan1_32 \
    python profiles/hbm_scan_probe.py 10000000 > $out/dev_ncu_scan1_32.log 2>&1
ls -la $out | grep dev_
