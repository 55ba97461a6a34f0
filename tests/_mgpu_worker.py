"""torchrun worker for tests/test_multigpu.py: a view-sharded optimizer step on G GPUs must equal -- BIT FOR BIT -- a
single-GPU Adam step on the per-rank gradients summed in rank order.

The blend backward accumulates with floating-point atomics, so re-rendering a view reproduces its gradient only up
to the last bits (and Adam turns a last-bit difference of a near-zero gradient into a full step of either sign).  The
comparison therefore uses the gradients each rank ACTUALLY produced (captured before the exchange and gathered on
rank 0): with them the expected parameters are fully determined and no outlier allowance is needed."""
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
graph = mode.endswith("+graph")          # the whole iteration incl. the exchange replayed from CUDA graphs
mode = mode.replace("+graph", "")
N_STEPS = 7 if graph else 3              # graph: step 0 eager (sizes the binning), then capture + replays per view
tr = I.JointTrainer(sc, dev, gt_images=gts, world_size=world, rank=rank, exchange=mode, use_graph=graph)
assert tr.use_graph == graph
captured = []
if graph:
    last_g, last_pg = torch.zeros_like(tr.grads), torch.zeros_like(tr.pose_grad)
    tr._after_backward = lambda: (last_g.copy_(tr.grads), last_pg.copy_(tr.pose_grad))
for s in range(N_STEPS):
    v = view_for_step(sc.n_views, world, rank, s)
    if graph:
        tr.step(v)
        captured.append((last_g.clone(), last_pg.clone()))
        continue
    tr.iteration += world
    tr._launch_forward(v)
    tr.loss_and_backward(v, tr.gt[v])
    captured.append((tr.grads.clone(), tr.pose_grad.clone()))
    if tr._fused:
        tr.fused_exchange_step()
    else:
        tr.reduce_grads()
        tr.optimizer_step()
    assert tr._settle()
torch.cuda.synchronize()
tr.check_peer_errors()
if graph:
    assert len(tr._graphs) == len(set(view_for_step(sc.n_views, world, rank, s) for s in range(N_STEPS))), tr._graphs.keys()
# every replica must hold identical parameters
mine = tr.params.clone()
ref0 = mine.clone()
dist.broadcast(ref0, src=0)
assert torch.equal(mine, ref0), f"rank {rank}: replicas diverged"
poses0 = tr.poses.clone()
dist.broadcast(poses0, src=0)
assert torch.equal(tr.poses, poses0), f"rank {rank}: pose tables diverged"
# gather the per-rank gradients of every step on rank 0
gathered = []
for g, pg in captured:
    gl = [torch.empty_like(g) for _ in range(world)]
    pl = [torch.empty_like(pg) for _ in range(world)]
    dist.all_gather(gl, g.contiguous())
    dist.all_gather(pl, pg.contiguous())
    gathered.append((gl, pl))
if rank == 0:
    one = I.JointTrainer(sc, dev, gt_images=gts)
    for s, (gl, pl) in enumerate(gathered):
        acc, pacc = gl[0].clone(), pl[0].clone()
        for r in range(1, world):
            acc += gl[r]
            pacc += pl[r]
        one.grads.copy_(acc)
        one.pose_grad.copy_(pacc)
        one.iteration += world
        one.optimizer_step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    diff = float((one.params - mine).abs().max())
    perr = float((one.poses - tr.poses).abs().max())
    # NCCL may add in a different order for more than two ranks; the fused kernel adds in rank order for any world
    tol = 0.0 if (tr._fused or world == 2) else 1e-6
    assert diff <= tol and perr <= tol, (mode, diff, perr)
    # the step must actually have moved the parameters
    moved = float((one.view(one.params, "xyz") - sc.params["xyz"].to(dev)).abs().max())
    assert moved > 0
    print(f"MGPU_OK mode={mode}{'+graph' if graph else ''} world={world} steps={N_STEPS} max param diff {diff:.1e} pose diff {perr:.1e}")
dist.barrier()
tr.close()
dist.destroy_process_group()
