set -x
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.err
timeout 800 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_k20.json 2> gpurun_out/r02_bench_reference_k20.err; tail -c 200 gpurun_out/r02_bench_reference_k20.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_final.json"))
print("final", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "e2e_trainer", d["e2e_trainer"]["value"], "steps", d["steps"], "launches", d["gpu_launches"])
print({k:round(v["ms"],4) for k,v in d["kernels"].items() if isinstance(v,dict) and "ms" in v})
print("dropin", {k:round(v["value"],1) for k,v in d["dropin_boundary"].items()}, "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["spread"], d["clocks"])
r=json.load(open("gpurun_out/r02_bench_reference_k20.json")); print("reference_k20", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["spread"])
PY
