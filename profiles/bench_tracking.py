"""Benchmark of the pose-only tracking mode (/root/reference/render.py:99-170: 500 Adam
iterations per test view on the 7 pose parameters, Gaussians frozen) on the headline scene (1M Gaussians, 1920x1080,
SH degree 3).  Prints one JSON line: ms per tracking iteration (device-resident loop, no host sync inside), the time for
the reference's 500 iterations of one view, and the same iteration done the training way (full backward through the
autograd boundary with only the pose requiring grad) for comparison.
    python profiles/bench_tracking.py [--iters 200] [--config 2]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantsplat_b200 as I
from instantsplat_b200.scenes import make_config
from instantsplat_b200.tracking import PoseTracker

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sc = make_config(a.config, a.scale)
p = sc.params
tr = PoseTracker(p["xyz"], p["rotation"], p["scaling"], p["opacity"], p["f_dc"], p["f_rest"], sc.width, sc.height, sc.fovx,
                 sc.fovy, sh_degree=sc.sh_degree, device=dev)
view = 5
gt = tr.render(sc.poses[view])
init = sc.poses[view] + torch.tensor([0.0, 0.004, -0.003, 0.002, 0.02, -0.015, 0.02])
tr.optimize(init, gt, num_iter=10)                       # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pose, best, trace = tr.optimize(init, gt, num_iter=a.iters, return_trace=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
# the same iteration through the generic autograd boundary (render() + torch masked L1 + backward + torch Adam)
from instantsplat_b200.camera import SimpleCamera
cam = SimpleCamera(sc.width, sc.height, sc.fovx, sc.fovy, device=dev)
pc = I.GaussianModel.from_scene(sc, dev)
for t in (pc._xyz, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling, pc._rotation):
    t.requires_grad_(False)
q = init[:4].to(dev).clone().requires_grad_()
T = init[4:].to(dev).clone().requires_grad_()
opt = torch.optim.Adam([{"params": [T], "lr": 0.003}, {"params": [q], "lr": 0.001}], betas=(0.9, 0.999), weight_decay=1e-4)
pipe, bg = I.PipelineDefaults(), torch.zeros(3, device=dev)

def generic_iter():
    img = I.render(cam, pc, pipe, bg, camera_pose=torch.cat([q, T]))["render"]
    mask = (img > 0.0).float()
    loss = ((img - gt).abs() * mask).sum() / mask.sum()
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)

for _ in range(5):
    generic_iter()
torch.cuda.synchronize()
n2 = min(a.iters, 50)
e0.record()
for _ in range(n2):
    generic_iter()
e1.record()
torch.cuda.synchronize()
ms_gen = e0.elapsed_time(e1) / n2
print(json.dumps({
    "metric": "tracking_ms_per_iteration", "value": ms, "unit": "ms", "iters_timed": a.iters,
    "workload": f"{sc.P} Gaussians, {sc.width}x{sc.height}, SH deg {sc.sh_degree}, one test view, pose-only (render.py:99-170)",
    "seconds_per_view_500_iters": ms * 500 / 1e3,
    "loss_first_last": [float(trace[0]), float(trace[-1])], "best_loss": best,
    "generic_autograd_path_ms_per_iteration": ms_gen,
    "generic_path": "render() [fused, only the pose requires grad] + torch masked L1 + backward + torch.optim.Adam",
    "instances": tr.R_seen}))
