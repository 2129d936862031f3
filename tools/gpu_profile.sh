# Profiling recipe of this repo (run under gpurun).  1) launch list + DRAM traffic of the bench command,
# 2) full-set capture of the sweep kernel in the stationary regime on a shorter variant (p=100), binomial and gaussian.
mkdir -p gpurun_out
unset CGG_PROFILE
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_cfg3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_cfg3.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/plain_cfg3.log | cut -c1-200
CMD2="python bench.py --workload cfg3 --cols 100 --steps 1 --warmup 3 --burnin-iters 30 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain_p100.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 4 -c 1 -f -o gpurun_out/prof_binom_jet_r01 $CMD2 > gpurun_out/ncu_full.log 2>&1
CMD3="python bench.py --workload cfg3 --cols 100 --family gaussian --steps 1 --warmup 3 --burnin-iters 30 --no-e2e --no-cpu"
$CMD3 > gpurun_out/plain_g100.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 4 -c 1 -f -o gpurun_out/prof_gauss_jet_r01 $CMD3 > gpurun_out/ncu_full_g.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_cfg3.csv
# the other workloads, plain
for w in "--workload cfg2" "--workload cfg4" "--workload cfg3 --family gaussian --cols 100" "--workload cfg3 --no-jet" "--workload cfg4 --no-jet" "--workload cfg2 --no-jet"; do
  echo "== $w"; timeout 600 python bench.py $w --steps 3 --warmup 3 --no-cpu --e2e-iters 100 2>&1 | tail -1
done > gpurun_out/bench_others.log 2>&1
