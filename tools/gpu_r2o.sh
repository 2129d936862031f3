mkdir -p gpurun_out
( for q in 161 1 1 1 1; do
  echo "== binomial quad=$q"; CGG_QUAD=$q CGG_SMALLN=0 timeout 60 python tools/quad_debug.py binomial 150001 8 2 2>&1 | grep -v "^  File\|^    \|^Traceback" | tail -1 | cut -c1-300
done
for q in 1 1; do echo "== gaussian quad=$q C=12"; CGG_QUAD=$q CGG_SMALLN=0 timeout 60 python tools/quad_debug.py gaussian 150001 12 3 2>&1 | grep -v "^  File\|^    \|^Traceback" | tail -1 | cut -c1-300; done
) > gpurun_out/r2o.log 2>&1
cat gpurun_out/r2o.log
( timeout 900 python -m pytest tests/test_gpu_edges.py -x -q -k "group_passes or handover_stress or pair_passes" 2>&1 | tail -5 ) > gpurun_out/r2n_tests.log 2>&1
cat gpurun_out/r2n_tests.log
