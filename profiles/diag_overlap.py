"""Overlap-mode diagnostics: phase timeline (CUDA events on both streams) and step time vs the side-stream Adam's grid."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import instantsplat_b200 as I
from instantsplat_b200.scenes import make_config, perturbed_copy

dev = torch.device("cuda", 0)
sc = make_config(2, 1.0)
tgt = I.JointTrainer(sc, dev)
pp = perturbed_copy(sc, sigma=0.05)
for k, kk in (("xyz", 3), ("f_dc", 3), ("opacity", 1), ("scaling", 3)):
    tgt.view(tgt.params, k).copy_(pp[k].reshape(sc.P, kk).to(dev))
gt = torch.stack([tgt.render(v).clone() for v in range(sc.n_views)])
del tgt
out = {}
for name, kw in (("serial", dict(overlap=False)), ("ov_296", dict(overlap=True, ctas=296)), ("ov_592", dict(overlap=True, ctas=592)),
                 ("ov_1184", dict(overlap=True, ctas=1184)), ("ov_148", dict(overlap=True, ctas=148))):
    tr = I.JointTrainer(sc, dev, gt_images=gt, use_graph=True, overlap=kw["overlap"])
    if kw["overlap"]:
        tr.overlap_ctas = kw["ctas"]
    for s in range(30):
        tr.step(s % sc.n_views)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(30, 130):
        tr.step(s % sc.n_views)
    tr.join()
    e1.record()
    torch.cuda.synchronize()
    rec = {"ms_per_step": e0.elapsed_time(e1) / 100}
    if kw["overlap"]:
        tr._timeline = []
        for s in range(130, 170):
            tr.step(s % sc.n_views)
        torch.cuda.synchronize()
        tl = tr._timeline[5:]
        avg = lambda f: sum(f(e) for e in tl) / len(tl)
        rec["timeline_ms_from_step_start"] = {
            "color_done(side)": avg(lambda e: e[0].elapsed_time(e[1])),
            "graphA_done(main)": avg(lambda e: e[0].elapsed_time(e[2])),
            "graphB_done(main)": avg(lambda e: e[0].elapsed_time(e[3])),
            "sh_adam_start(side)": avg(lambda e: e[0].elapsed_time(e[4])),
            "sh_adam_done(side)": avg(lambda e: e[0].elapsed_time(e[5])),
        }
        # previous step's SH Adam end relative to THIS step's start
        rec["prev_sh_adam_done_after_step_start_ms"] = sum(tl[i][0].elapsed_time(tl[i - 1][5]) for i in range(1, len(tl))) / (len(tl) - 1)
        tr._timeline = None
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
    del tr
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join("gpurun_out", "r02_diag_overlap.json"), "w"), indent=1)
