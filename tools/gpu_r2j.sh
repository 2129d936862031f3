mkdir -p gpurun_out
python tools/transient_probe.py --iters 40 > gpurun_out/r2j_transient.log 2>&1
cat gpurun_out/r2j_transient.log | cut -c1-400
