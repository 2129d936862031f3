# The table of BASELINE.md section 4 (one GPU): every single-GPU configuration, exact-pass rows for K = 1, 2, 4, 8, tuned w.
R=${R:-r02}
mkdir -p gpurun_out
one() { echo "== $1"; shift; timeout 900 "$@" 2>&1 | tail -1; }
( one "cfg3" python bench.py --steps 5 --warmup 3
  one "cfg2" python bench.py --workload cfg2 --steps 5 --warmup 3
  one "cfg4" python bench.py --workload cfg4 --steps 5 --warmup 3
  one "cfg1 (README shape: gaussian n=1000 p=3, 1 chain)" python bench.py --workload cfg2 --rows 1000 --cols 3 --family gaussian --chains 1 --steps 200 --warmup 3 --no-cpu --e2e-iters 500
  one "gaussian n=1e6 p=100" python bench.py --family gaussian --cols 100 --steps 5 --warmup 3 --no-cpu
  one "cfg3 tuned w = 4/sqrt(n)" python bench.py --w 0.004 --steps 3 --warmup 3 --no-cpu --no-e2e
  for K in 1 2 4 8; do one "cfg3 p=100 exact passes only, K=$K" python bench.py --cols 100 --no-jet --K $K --steps 2 --warmup 3 --no-cpu --no-e2e; done
  for K in 1 8; do one "cfg4 p=100 exact passes only, K=$K" python bench.py --workload cfg4 --cols 100 --no-jet --K $K --steps 2 --warmup 3 --no-cpu --no-e2e; done
  one "reference arm (CPU port)" python bench.py --impl reference --steps 2 --warmup 1
) > gpurun_out/${R}_matrix.log 2>&1
python - <<'PY'
import json, os
R = os.environ.get("R", "r02")
for line in open(f"gpurun_out/{R}_matrix.log"):
    if line.startswith("=="): print(line.strip()); continue
    try: d = json.loads(line)
    except Exception: print("   ", line[:200].strip()); continue
    e2e = d.get("e2e") or {}
    rf = d.get("roofline") or {}
    cb = d.get("cpu_baseline") or {}
    print("    value %.0f  e2e %s  frac %s  achieved %s GB/s  l2alg %s GB/s  cpu %s (%s cores)  passes/update %s" % (
        d["value"], e2e.get("value") and round(e2e["value"]), rf.get("frac") and round(rf["frac"], 3), rf.get("achieved") and round(rf["achieved"]),
        rf.get("l2_algorithmic_gbs") and round(rf["l2_algorithmic_gbs"]), cb.get("value") and round(cb["value"], 1), cb.get("cores"),
        d.get("engine_stats", {}).get("chain_passes_per_update")))
PY
