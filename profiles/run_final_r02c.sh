#!/bin/bash
# Round-2 closing verification on ONE GPU with the final tree: GPU test suite, smoke(), the headline bench line, config 5
# and the product-shaped run.  From the repo root through gpurun; results in gpurun_out/ (copied by hand to profiles/).
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -q -m gpu > $out/r02c_gputests.log 2>&1; tail -2 $out/r02c_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $out/r02c_smoke.log 2>&1; tail -1 $out/r02c_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > $out/r02c_bench_N1.json 2> $out/r02c_bench_N1.err
timeout 200 python bench.py --config c5 --steps 5 --warmup 2 > $out/r02c_bench_c5.json 2> $out/r02c_bench_c5.err
for f in r02c_bench_N1 r02c_bench_c5; do
  python - <<PY
import json
try:
    d = json.loads(open("$out/$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("ids_match_gpu", (d.get("cpu_baseline") or {}).get("codes_match_gpu")))
except Exception as e:
    print("$f FAILED", e)
PY
done
