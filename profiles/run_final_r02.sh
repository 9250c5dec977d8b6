#!/bin/bash
# Round-2 closing run on ONE GPU (from the repo root, through gpurun): the GPU test suite, smoke(), and the plain bench
# lines of every single-GPU configuration with the final code.  profiles/collect_r02.sh copies the results into profiles/.
out=gpurun_out; mkdir -p $out
python -m pytest tests -q -m gpu > $out/r02_gputests.log 2>&1; tail -2 $out/r02_gputests.log
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $out/r02_smoke.log 2>&1; tail -1 $out/r02_smoke.log
python bench.py --steps 20 --warmup 3 > $out/bench_r02e_N1.json 2> $out/bench_r02e_N1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_r02_ref.json 2> $out/bench_r02_ref.err
python bench.py --config c1 > $out/bench_r02_c1.json 2> $out/bench_r02_c1.err
python bench.py --config c2 > $out/bench_r02_c2.json 2> $out/bench_r02_c2.err
python bench.py --config c3 --steps 10 --warmup 3 > $out/bench_r02_c3.json 2> $out/bench_r02_c3.err
python bench.py --config c5 --steps 5 --warmup 2 > $out/bench_r02_c5.json 2> $out/bench_r02_c5.err
python bench.py --config pv > $out/bench_r02_pv.json 2> $out/bench_r02_pv.err
python profiles/hbm_scan_probe.py > $out/hbm_probe_r02h.json 2> $out/hbm_probe_r02h.err
for f in bench_r02e_N1 bench_r02_ref bench_r02_c1 bench_r02_c2 bench_r02_c3 bench_r02_c5 bench_r02_pv; do
  python - <<PY
import json
try:
    d = json.loads(open("$out/$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("unit"), "e2e", (d.get("e2e") or {}).get("value"), "recall", d.get("recall@10", (d.get("config") or {}).get("recall@10")))
except Exception as e:
    print("$f FAILED", e)
PY
done
