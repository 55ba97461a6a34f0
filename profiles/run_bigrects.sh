set -x
timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "big_rects or config2 or edge_cases or exact_cull or overflow or tracking or trainer_step or tiny_scene or properties" 2>&1 | tail -15
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c_bigrects.json 2> gpurun_out/r02_bench_c_bigrects.err; tail -c 400 gpurun_out/r02_bench_c_bigrects.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_c_bigrects.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_trainer"]["value"])
print({k:round(v["ms"],4) for k,v in d["kernels"].items() if isinstance(v,dict) and "ms" in v})
PY
