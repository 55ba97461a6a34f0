"""Diagnostic: per-kernel CUDA-event times of the forward kernels when (a) only forwards run back to back and (b) inside
full training steps -- shows how much of a kernel's in-step time is interference from its predecessor (e.g. the
optimizer's 1.6 GB of dirty lines still draining from L2 while the projection kernel issues its atomics)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantsplat_b200 as I
from instantsplat_b200 import _lib
from instantsplat_b200.scenes import make_config

sc = make_config(2)
gt = torch.rand(sc.n_views, 3, sc.height, sc.width, generator=torch.Generator().manual_seed(0)) * 0.5 + 0.25
tr = I.JointTrainer(sc, "cuda:0", gt_images=gt)
L = I.lib()
nk = len(_lib.KERNEL_IDS)

def collect():
    ms, cnt = (ctypes.c_double * nk)(), (ctypes.c_int64 * nk)()
    L.gsb_profile_collect(ms, cnt, nk)
    return {n: round(ms[i] / cnt[i], 4) for i, n in enumerate(_lib.KERNEL_IDS) if cnt[i]}

for s in range(5):
    tr.step(s % sc.n_views)
torch.cuda.synchronize()
out = {}
L.gsb_profile_enable(1)
for s in range(40):
    tr.render(s % sc.n_views)
torch.cuda.synchronize()
out["forward_only"] = collect()
for s in range(40):
    tr.step(s % sc.n_views)
torch.cuda.synchronize()
out["full_step_eager"] = collect()
# full steps, but with a device-wide pause between optimizer and next forward (lets L2 write back)
for s in range(40):
    tr.step(s % sc.n_views)
    torch.cuda.synchronize()
out["full_step_synced_each_step"] = collect()
L.gsb_profile_enable(0)
print(json.dumps(out))
