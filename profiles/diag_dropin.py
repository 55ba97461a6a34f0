"""Diagnosis aid: host wall time per phase of the drop-in (autograd) iteration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import instantsplat_b200 as I
from instantsplat_b200.camera import SimpleCamera
from instantsplat_b200.scenes import make_config
sc = make_config(2)
dev = torch.device("cuda:0")
gt = (torch.rand(sc.n_views, 3, sc.height, sc.width, generator=torch.Generator().manual_seed(0)) * 0.5 + 0.25).to(dev)
pc = I.SimpleGaussianModel(sc, dev); opt = pc.training_setup_pp()
cam = SimpleCamera(sc.width, sc.height, sc.fovx, sc.fovy, device=dev); pipe = I.PipelineDefaults(); bg = torch.zeros(3, device=dev)
acc = {}
def tick(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
for s in range(45):
    if s == 5: acc.clear()
    torch.cuda.synchronize(); t = time.perf_counter()
    v = s % sc.n_views
    pkg = I.render(cam, pc, pipe, bg, camera_pose=pc.get_RT(v)); t = tick("render", t)
    img = pkg["render"]
    loss = 0.8 * torch.abs(img - gt[v]).mean() + 0.2 * (1.0 - I.fused_ssim(img.unsqueeze(0), gt[v].unsqueeze(0))); t = tick("loss", t)
    loss.backward(); t = tick("backward", t)
    opt.step(); t = tick("opt.step", t)
    opt.zero_grad(set_to_none=True); t = tick("zero_grad", t)
print({k: round(1e3 * v / 40, 3) for k, v in acc.items()}, "ms per step (host wall incl. sync)")
print("mem allocated GB", torch.cuda.memory_allocated() / 1e9, "reserved", torch.cuda.memory_reserved() / 1e9, "num cudaMalloc", torch.cuda.memory_stats()["num_device_alloc"])
