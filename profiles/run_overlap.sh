set -x
timeout 600 python -X faulthandler -m pytest tests/test_reference_loop.py -x -q -m gpu -k "overlap or graph_replay" 2>&1 | tail -15
for o in "" "--no-overlap"; do
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline $o > gpurun_out/r02_bench_d_overlap$o.json 2> gpurun_out/r02_bench_d_overlap$o.err; tail -c 300 gpurun_out/r02_bench_d_overlap$o.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_d_overlap$o.json"))
print("$o", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_trainer"]["value"])
print({k:round(v["ms"],4) for k,v in d["kernels"].items() if isinstance(v,dict) and "ms" in v})
PY
done
