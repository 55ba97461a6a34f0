set -x
timeout 300 python -X faulthandler -m pytest tests/test_multigpu.py -x -q -m gpu -k "fused_p2p" 2>&1 | tail -15
for g in "--graph" "--no-graph"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 5 $g > gpurun_out/r02_n2_graph${g}.json 2> gpurun_out/r02_n2_graph${g}.err; tail -c 300 gpurun_out/r02_n2_graph${g}.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_n2_graph${g}.json"))
print("${g}", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["launch_mode"][:40])
PY
done
