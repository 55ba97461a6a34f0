"""torchrun worker for tests/test_multigpu.py: a 2-GPU view-sharded optimizer step must equal a
single-GPU step on the gradients accumulated over the same two views."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import instantsplat_b200 as I  # noqa: E402
from instantsplat_b200.parallel import init_from_env, view_for_step  # noqa: E402
from instantsplat_b200.scenes import surface_scene  # noqa: E402

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
sc = surface_scene(20_000, 4, 256, 192, seed=9, sh_degree=3)
gts = torch.rand(4, 3, sc.height, sc.width, generator=torch.Generator().manual_seed(2))
mode = os.environ.get("GSB_TEST_MODE", "allreduce")
tr = I.JointTrainer(sc, dev, gt_images=gts, world_size=world, rank=rank, exchange=mode)
for s in range(2):
    tr.step(view_for_step(sc.n_views, world, rank, s))
torch.cuda.synchronize()
# every replica must hold identical parameters
mine = tr.params.clone()
ref0 = mine.clone()
dist.broadcast(ref0, src=0)
assert torch.equal(mine, ref0), f"rank {rank}: replicas diverged"
if rank == 0:
    one = I.JointTrainer(sc, dev, gt_images=gts)
    for s in range(2):
        acc = torch.zeros_like(one.grads)
        pacc = torch.zeros_like(one.pose_grad)
        for r in range(world):
            v = view_for_step(sc.n_views, world, r, s)
            one.render(v)
            one.loss_and_backward(v, one.gt[v])
            acc += one.grads
            pacc += one.pose_grad
        one.grads.copy_(acc)
        one.pose_grad.copy_(pacc)
        one.iteration += 1
        one.optimizer_step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    # The blend backward accumulates with floating-point atomics, so two runs of the same view differ in the
    # last bits of the gradients.  Adam turns a gradient whose true value is ~0 (below that noise) into a
    # full-size step of random sign, so a handful of parameters may legitimately differ; everything else must
    # agree to rounding.
    diff = (one.params - mine).abs()
    frac_bad = float((diff > 2e-6).float().mean())
    perr = float((one.poses - tr.poses).abs().max())
    assert frac_bad < 2e-4 and perr < 1e-6, (frac_bad, float(diff.max()), perr)
    print(f"MGPU_OK mode={mode} outlier fraction {frac_bad:.2e} (max diff {float(diff.max()):.2e}) pose diff {perr:.2e}")
dist.barrier()
dist.destroy_process_group()
