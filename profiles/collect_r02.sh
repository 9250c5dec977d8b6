#!/bin/bash
# copy the round-2 artefacts that came back in gpurun_out/ (scratch) into profiles/ (tracked); run from the repo root
cp() { [ -f "$1" ] && command cp "$1" "$2" && echo "$2"; }
cp gpurun_out/bench_r02e_N1.json   profiles/r02_bench_N1.json
cp gpurun_out/bench_r02_ref.json    profiles/r02_bench_reference_arm.json
cp gpurun_out/hbm_probe_r02h.json  profiles/r02_hbm_scan.json
cp gpurun_out/bench_r02b_N2.json   profiles/r02_bench_N2.json
cp gpurun_out/bench_r02b_N4.json   profiles/r02_bench_N4.json
cp gpurun_out/bench_r02b_N8.json   profiles/r02_bench_N8.json
cp gpurun_out/bench_r02_c1.json    profiles/r02_bench_c1.json
cp gpurun_out/bench_r02_c2.json    profiles/r02_bench_c2.json
cp gpurun_out/bench_r02_c3.json    profiles/r02_bench_c3.json
cp gpurun_out/bench_r02_c3_N8.json profiles/r02_bench_c3_N8.json
cp gpurun_out/bench_r02_c5.json    profiles/r02_bench_c5.json
cp gpurun_out/bench_r02_c5_N8.json profiles/r02_bench_c5_N8.json
cp gpurun_out/bench_r02_pv.json    profiles/r02_bench_pv.json

cp gpurun_out/sweep_r02_c3.json    profiles/r02_c3_quota_sweep.json
cp gpurun_out/sweep_r02b.json      profiles/r02_c4_quota_sweep.json
cp gpurun_out/mp_check_r02c.log    profiles/r02_mp_sharded_check.log
cp gpurun_out/r02_gputests.log    profiles/r02_gputests.log
cp gpurun_out/r02_launches.csv     profiles/r02_launches.csv
cp gpurun_out/r02_sanitizer_memcheck.log  profiles/r02_sanitizer_memcheck.log
cp gpurun_out/r02_sanitizer_racecheck.log profiles/r02_sanitizer_racecheck.log
for t in scan_pk scan1 encode; do
  [ -f gpurun_out/r02_$t.ncu-rep ] && python profiles/ncu_summary.py gpurun_out/r02_$t.ncu-rep 30 > profiles/r02_${t}_ncu_full_summary.txt 2>/dev/null && echo profiles/r02_${t}_ncu_full_summary.txt
done
