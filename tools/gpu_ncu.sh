mkdir -p gpurun_out
unset CGG_PROFILE
CMD1="python bench.py --workload cfg3 --p 100 --family gaussian --chains 1 --steps 1 --warmup 1 --no-e2e --no-cpu"
CMD2="python bench.py --workload cfg3 --p 100 --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD1 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 1 -c 1 -f -o gpurun_out/prof_gauss_c1 $CMD1 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_persistent -s 1 -c 1 -f -o gpurun_out/prof_binom_c8 $CMD2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/*.ncu-rep
