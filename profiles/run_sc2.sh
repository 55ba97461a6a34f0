timeout 900 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny or config0 or config1_surface or edge_cases or big_rects or overflow or config2" 2>&1 | tail -3
cat > /tmp/pk.py <<'PY'
import json,sys
for v in sys.argv[1:]:
    d=json.load(open(f"gpurun_out/variant_{v}.json")); k=d["kernels"]
    print(v, {n:round(k[n]["ms"],4) for n in ("duplicate","sort_tile","scan")})
PY
bash profiles/run_variants.sh base sc2 base sc2; python /tmp/pk.py base sc2
