# round 2, run H: tests again (cluster driver, mailboxes), then diagnostics: what bounds the pair loop?
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2h_tests.log 2>&1
tail -12 gpurun_out/r2h_tests.log
B="python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 --workload cfg3 --cols 100"
: > gpurun_out/r2h_diag.log
for v in "-DCGG_DIAG_NOSTORE" "-DCGG_DIAG_NOMATH" "-DCGG_DIAG_NOSTORE -DCGG_DIAG_NOMATH"; do
  CGG_NVCC_EXTRA="$v" python -m mcmcglm_b200.build -f > /dev/null 2>&1
  echo "== [$v]" >> gpurun_out/r2h_diag.log
  timeout 300 $B 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/r2h_diag.log
  CGG_L2_PERSIST=0 timeout 300 $B 2>&1 | tail -1 | cut -c1-130 >> gpurun_out/r2h_diag.log
done
python -m mcmcglm_b200.build -f > /dev/null 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu 2>/dev/null && ./tools/fp64_probe > gpurun_out/r2h_fp64_probe.json 2>&1
cat gpurun_out/r2h_diag.log; cat gpurun_out/r2h_fp64_probe.json
