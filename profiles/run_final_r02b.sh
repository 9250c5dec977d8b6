#!/bin/bash
# Round-2 closing run, second session (tensor-core fine argmin, streamed coarse assignment, pipelined host encode): the GPU
# test suite, smoke(), the bench lines the new encode kernels change, a launch list and --set full captures of the encode
# kernels, memcheck of smoke().  From the repo root through gpurun; results land in gpurun_out/ (copied by hand to profiles/).
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -q -m gpu > $out/r02b_gputests.log 2>&1; tail -2 $out/r02b_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $out/r02b_smoke.log 2>&1; tail -1 $out/r02b_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > $out/r02b_bench_N1.json 2> $out/r02b_bench_N1.err
timeout 200 python bench.py --config c5 --steps 5 --warmup 2 > $out/r02b_bench_c5.json 2> $out/r02b_bench_c5.err
timeout 300 python bench.py --config pv > $out/r02b_bench_pv.json 2> $out/r02b_bench_pv.err
timeout 200 python bench.py --config c2 > $out/r02b_bench_c2.json 2> $out/r02b_bench_c2.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/r02b_launches_encode.csv \
    python profiles/dev/ftc_probe.py --time-only > $out/r02b_launches_encode.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_fine_tc|k_coarse_big|k_fine_redo|k_rotate_dmma" -s 8 -c 4 -f \
    -o $out/r02b_encode python profiles/dev/ftc_probe.py --time-only > $out/r02b_ncu_encode.log 2>&1; tail -2 $out/r02b_ncu_encode.log
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $out/r02b_sanitizer_memcheck.log 2>&1; tail -3 $out/r02b_sanitizer_memcheck.log
for f in r02b_bench_N1 r02b_bench_c5 r02b_bench_pv r02b_bench_c2; do
  python - <<PY
import json
try:
    d = json.loads(open("$out/$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), "encode", (d.get("encode") or {}).get("codes_per_s"))
except Exception as e:
    print("$f FAILED", e)
PY
done
