"""Diagnostic: how the instance count and the radius distribution evolve while the headline scene trains (explains
why the binning kernels slow down over a long bench run)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantsplat_b200 as I
from instantsplat_b200.scenes import make_config, perturbed_copy

sc = make_config(2)
dev = "cuda:0"
tgt = I.JointTrainer(sc, dev)
pp = perturbed_copy(sc, sigma=0.05)
for k, kk in (("xyz", 3), ("f_dc", 3), ("opacity", 1), ("scaling", 3)):
    tgt.view(tgt.params, k).copy_(pp[k].reshape(sc.P, kk).to(dev))
gt = torch.stack([tgt.render(v).clone() for v in range(sc.n_views)])
del tgt
import ctypes
from instantsplat_b200 import _lib
L = I.lib()
nk = len(_lib.KERNEL_IDS)
tr = I.JointTrainer(sc, dev, gt_images=gt)
rows = []
L.gsb_profile_enable(1)
for s in range(0, 241):
    tr.step(s % sc.n_views)
    if s % 40 == 0:
        torch.cuda.synchronize()
        ms, cnt = (ctypes.c_double * nk)(), (ctypes.c_int64 * nk)()
        L.gsb_profile_collect(ms, cnt, nk)
        kern = {n: round(ms[i] / cnt[i], 4) for i, n in enumerate(_lib.KERNEL_IDS) if cnt[i]}
        r = tr.radii.float()
        st = tr._status_t.cpu().tolist()
        sca = torch.exp(tr.view(tr.params, "scaling")).max(dim=1).values
        rows.append(dict(step=s, kernels_ms_avg_since_last_row=kern, R=tr.last_R, longest_list=st[2], n_large=st[4], n_huge=st[5],
                         radius_mean=float(r.mean()), radius_p99=float(r.quantile(0.99)), radius_max=float(r.max()),
                         n_radius_gt_64=int((r > 64).sum()), n_radius_gt_256=int((r > 256).sum()),
                         n_radius_gt_1024=int((r > 1024).sum()), scale_max=float(sca.max()), scale_p999=float(sca.quantile(0.999)),
                         opacity_mean=float(torch.sigmoid(tr.view(tr.params, "opacity")).mean())))
print(json.dumps(rows))
