# Round-2 final profiling pass (one GPU): launch list + DRAM traffic of the exact bench command, --set full captures of the
# sweep kernel (binomial, steady state) and of the cluster kernel at cfg2; plus two A/B lines that ride along.
R=${R:-r02b}
mkdir -p gpurun_out
unset CGG_PROFILE
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/${R}_plain_cfg3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_launches_cfg3.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
tail -1 gpurun_out/${R}_plain_cfg3.log | cut -c1-160
P100="python bench.py --workload cfg3 --cols 100 --steps 1 --warmup 3 --burnin-iters 30 --no-e2e --no-cpu"
$P100 > gpurun_out/${R}_plain_p100.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 4 -c 1 -f -o gpurun_out/${R}_sweep_binomial $P100 > gpurun_out/${R}_ncu_full.log 2>&1
C2="python bench.py --workload cfg2 --steps 3 --warmup 3 --no-e2e --no-cpu"
$C2 > gpurun_out/${R}_plain_cfg2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_cluster -s 4 -c 1 -f -o gpurun_out/${R}_cluster_cfg2 $C2 > gpurun_out/${R}_ncu_full_c2.log 2>&1
ls -la gpurun_out/${R}_*.ncu-rep gpurun_out/${R}_launches_cfg3.csv
