"""Profiling driver: a few JointTrainer steps on the headline workload with random GT images, so that
`ncu` sees only hot-path launches.  Usage (under gpurun, see B200_PROFILING.md):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python profiles/prof_step.py --steps 3
  ncu --set full --clock-control none --import-source on -k regex:k_blend_bwd -s 2 -c 1 -o gpurun_out/blend_bwd \
      python profiles/prof_step.py --steps 3
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantsplat_b200 as I
from instantsplat_b200.scenes import make_config

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
sc = make_config(a.config, a.scale)
gt = torch.rand(sc.n_views, 3, sc.height, sc.width, generator=torch.Generator().manual_seed(0)) * 0.5 + 0.25
tr = I.JointTrainer(sc, "cuda:0", gt_images=gt)
for s in range(a.steps):
    tr.step(s % sc.n_views)
torch.cuda.synchronize()
print("R", tr.last_R, "loss", float(tr.loss_value()))
