#!/bin/bash
# Round evidence on the GPU box: tests, smoke, bench lines, launch list, ncu captures exported as raw CSV (the .ncu-rep files
# are deleted: gpurun copies back at most 64 MiB).   usage: tools/evidence.sh <tag>
T=${1:-rX}
O=gpurun_out
mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/${T}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1
python bench.py > $O/${T}_bench_uniform_n1.json 2> $O/${T}_e1.log
python bench.py --state clustered > $O/${T}_bench_clustered_n1.json 2> $O/${T}_e2.log
python bench.py --state clumpy > $O/${T}_bench_clumpy_n1.json 2> $O/${T}_e3.log
python bench.py --impl reference > $O/${T}_bench_reference.json 2> $O/${T}_e4.log
python bench.py --np-side 512 --steps 2 --warmup 3 --subcycle 5 > $O/${T}_bench_uniform_512.json 2> $O/${T}_e5.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_b.log 2>&1
cap() {  # name, kernel regex, skip, count, extra bench args
  timeout 400 ncu --set full --clock-control none -k "regex:$2" -s $3 -c $4 -o $O/$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline $5 > $O/$1.log 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  rm -f $O/$1.ncu-rep
}
cap ${T}_force '^k_force' 0 2 ""
cap ${T}_force_clustered '^k_force' 0 2 "--state clustered"
cap ${T}_build 'k_scatter|k_cm_tile|k_left_count|k_gather' 40 9 ""
cat $O/${T}_pytest_gpu.log $O/${T}_smoke.log
tail -n 2 $O/${T}_e1.log $O/${T}_e3.log $O/${T}_e5.log
du -sh $O
