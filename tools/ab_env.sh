# usage: bash tools/ab_env.sh VAR v1 v2 ...   -- headline bench once per value of the environment variable VAR
VAR=$1; shift
for h in "$@"; do
env $VAR=$h python bench.py --steps 5 --warmup 3 --no-subrecords --no-gpu-baseline --no-cpu-baseline > gpurun_out/ab_$h.json 2>gpurun_out/ab_$h.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$h.json"))
print("$VAR=$h", round(d["value"]/1e6,1), "M/s e2e", round(d["e2e"]["value"]/1e6,1), " ".join(f"{k}={v['avg_ms']:.4f}" for k,v in d["kernels"].items()))
PY
done
