# A/B of library builds: instantsplat_b200/lib/variants/<name>.so copied over the in-tree library one at a time
cp instantsplat_b200/lib/libgsb200.so /tmp/keep.so
for v in "$@"; do
  cp instantsplat_b200/lib/variants/$v.so instantsplat_b200/lib/libgsb200.so
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err || tail -5 gpurun_out/variant_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/variant_$v.json"))
k=d["kernels"]
print("$v", round(d["value"],1), round(d["ms_per_step"],4), {n:round(k[n]["ms"],4) for n in ("preprocess","preprocess_bwd","loss_fwd","loss_bwd","blend_bwd","blend_fwd")})
PY
done
cp /tmp/keep.so instantsplat_b200/lib/libgsb200.so
