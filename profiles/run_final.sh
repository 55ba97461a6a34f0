# Final validation of a build: full GPU suite, smoke, default bench line, and both arms the way the driver launches them.
set -x
timeout 1200 python -X faulthandler -m pytest tests -q -m gpu 2>&1 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final_k20.json 2> gpurun_out/r02_bench_final_k20.err; tail -c 200 gpurun_out/r02_bench_final_k20.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_k20.json 2> gpurun_out/r02_bench_reference_k20.err; tail -c 200 gpurun_out/r02_bench_reference_k20.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_final_k20.json"))
print("final_k20", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "e2e_trainer", d["e2e_trainer"]["value"], "steps", d["steps"], "launches", d["gpu_launches"])
print({k:round(v["ms"],4) for k,v in d["kernels"].items() if isinstance(v,dict) and "ms" in v})
print("dropin", {k:round(v["value"],1) for k,v in d["dropin_boundary"].items()}, "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["spread"])
r=json.load(open("gpurun_out/r02_bench_reference_k20.json")); print("reference_k20", r["value"], r["cpu_baseline"]["sample"][:260], r["cpu_baseline"]["spread"])
PY
