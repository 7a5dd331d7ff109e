run() { timeout 200 python bench.py --no-cpu-baseline --no-fp32 --no-latency --no-extra --steps 10 --warmup 3 > gpurun_out/quick_$1.json 2> gpurun_out/quick_$1.err; tail -3 gpurun_out/quick_$1.err; python - <<PY
import json
d = json.loads(open("gpurun_out/quick_$1.json").read().strip().splitlines()[-1])
print("$1 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
print(" ".join("%s=%.2f" % (s["stage"].replace("gfe.bottleneck","b"), s["us_per_image"]) for s in d.get("stages", [])))
PY
}
