#!/bin/bash
# ncu evidence of one round (B200_PROFILING.md recipe): launch list of the bench command, then one full capture each of
# the traversal and shade kernels.  usage: tools/profile.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r2}
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-also"
$cmd > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_list.log 2>&1
for k in extend shade shadow; do
  ncu --set full --clock-control none --import-source on -k regex:k_wf_${k} -s 32 -c 2 -o gpurun_out/${tag}_${k} -f $cmd > gpurun_out/${tag}_ncu_${k}.log 2>&1
done
ls -la gpurun_out/${tag}_*
