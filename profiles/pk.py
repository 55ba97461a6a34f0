import json, sys
for v in sys.argv[1:]:
    d = json.load(open(f"gpurun_out/variant_{v}.json")); k = d["kernels"]
    print(v, {n: round(k[n]["ms"], 4) for n in ("duplicate", "sort_tile", "scan")})
