mkdir -p gpurun_out/r2n
timeout 300 python -m pytest tests/test_gpu_bf16.py -x -q > gpurun_out/r2n/pytest_bf16_$1.log 2>&1; tail -1 gpurun_out/r2n/pytest_bf16_$1.log
timeout 300 python bench.py --no-train --no-eager --no-cpu-baseline --no-extra --no-fp32 --no-latency > gpurun_out/r2n/bench_$1.json 2> gpurun_out/r2n/bench_$1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2n/bench_$1.json"))
print("$1 value", round(d["value"]), "full", round(d["full_step_us_per_image_one_launch_set"], 2))
print(" ".join("%s=%.2f" % (s["stage"].replace("gfe.bottleneck", "b"), s["us_per_image"]) for s in d["stages"]))
PY
