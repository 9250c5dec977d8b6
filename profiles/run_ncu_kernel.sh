#!/bin/bash
# one --set full capture of the kernels matching $2 (regex) from a short bench run; tag = $1
tag=$1; pat=$2; skip=${3:-4}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 2 -f -o gpurun_out/${tag} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${tag}.log 2>&1
tail -c 600 gpurun_out/ncu_${tag}.log
