#!/bin/bash
# ncu captures of the main kernels, summarised ON THE GPU BOX (the .ncu-rep files with imported
# source are 15-40 MB each; gpurun brings back at most 64 MiB).  Writes gpurun_out/r02_*_ncu_*.txt
# and gpurun_out/r02_traffic_raw.txt.  Run under gpurun from the repository root.
T=profiles/tools
O=gpurun_out
mkdir -p $O
cap() {  # name kernel-regex skip command...
  local name=$1 regex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 \
      -o /tmp/$name "$@" > $O/ncu_$name.log 2>&1
  python $T/ncu_summary.py /tmp/$name.ncu-rep > $O/${name}_ncu_summary.txt 2>&1
  python $T/ncu_lines.py /tmp/$name.ncu-rep 40 > $O/${name}_ncu_lines.txt 2>&1
  python $T/ncu_traffic.py $name=/tmp/$name.ncu-rep >> $O/r02_traffic_raw.txt 2>&1
}
cap r02_c1_dmma leapfrog_dmma 3 python bench.py --no-cpu-baseline --no-workloads --steps 2 --warmup 1
cap r02_c2_softabs implicit_leapfrog 0 python $T/run_cfg.py C2 1 1
cap r02_c6_softabs_dense implicit_leapfrog 0 python $T/run_cfg.py C6 1 1 296
cap r02_c3_constrained constrained_ 0 python $T/run_cfg.py C3 50 1
cap r02_c4_dense implicit_leapfrog 0 python $T/run_cfg.py C4 1 1 148
cap r02_nuts_c1 nuts_dmma 1 python $T/run_nuts.py 8192 6 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/r02_bench_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 1 \
    > $O/ncu_launches.log 2>&1
cp profiles/r02_traffic.json $O/r02_traffic.json 2>/dev/null
ls -la $O | tail -30
