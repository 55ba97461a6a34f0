"""fp64 central-difference check of the CPU oracle's rasterizer (CPU test, no GPU).

The rasterizer core of the oracle (EWA projection, conic, SH colour, depth-ordered alpha blend, the InstantSplat
pose pre-transform) cannot be pinned to the reference's own source -- the upstream submodule is empty ("parity
unpinned") -- so the one independent check available is mathematical: on a 64-Gaussian scene
in fp64, the oracle's autograd gradient of sum(w * image) must equal central finite differences for EVERY input
(xyz, raw quaternion, log-scale, opacity logit, SH dc / rest, and the 7 pose parameters).

The comparison runs in the smooth regime, where upstream's analytic backward coincides
with the true derivative: alpha stays below the 0.99 clamp, no Gaussian hits the 1.3*tan(fov) frustum clamp, and
the alpha >= 1/255 / T < 1e-4 cut-offs (jump discontinuities of the loss) are switched off through the oracle's
`alpha_min` / `t_eps` arguments.  The remaining deliberate deviation, 1/(det^2 + 1e-7) in the conic backward, is
below 2e-6 relative here and inside the tolerance.
"""
import math

import torch

from oracle import gs_oracle as O

TOL = 2e-5


def _scene(seed=3, P=64, W=64, H=48):
    g = torch.Generator().manual_seed(seed)
    dd = torch.float64
    xyz = torch.rand(P, 3, generator=g, dtype=dd)
    xyz[:, 0:2] = xyz[:, 0:2] * 1.6 - 0.8
    xyz[:, 2] = xyz[:, 2] * 2.5 + 2.5
    prm = dict(xyz=xyz,
               rotation=torch.randn(P, 4, generator=g, dtype=dd),
               scaling=math.log(0.12) + 0.3 * torch.randn(P, 3, generator=g, dtype=dd),
               opacity=torch.randn(P, 1, generator=g, dtype=dd).clamp(-2.0, 1.5),
               f_dc=torch.randn(P, 1, 3, generator=g, dtype=dd),
               f_rest=0.2 * torch.randn(P, 15, 3, generator=g, dtype=dd))
    pose = torch.tensor([0.97, 0.05, -0.08, 0.03, 0.04, -0.03, 0.1], dtype=dd)
    fovx = math.radians(60.0)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
    cam = O.Camera.instantsplat(W, H, fovx, fovy, bg=torch.tensor([0.1, 0.2, 0.3], dtype=dd), sh_degree=3, dtype=dd)
    w = torch.rand(3, H, W, generator=g, dtype=dd)
    return prm, pose, cam, w


def _loss(prm, pose, cam, w):
    img, _ = O.render_instantsplat(prm["xyz"], prm["rotation"], prm["scaling"], prm["opacity"], prm["f_dc"],
                                   prm["f_rest"], pose, cam, alpha_min=0.0, t_eps=0.0)
    return (img * w).sum()


def test_oracle_rasterizer_gradients_match_fp64_central_differences():
    prm, pose, cam, w = _scene()
    leaves = {k: v.clone().requires_grad_(True) for k, v in prm.items()}
    pose_l = pose.clone().requires_grad_(True)
    # the smooth regime really holds for this scene
    with torch.no_grad():
        means, rots = O.pose_pretransform(prm["xyz"], prm["rotation"], pose)
        pr = O.project(means, torch.exp(prm["scaling"]), rots, torch.sigmoid(prm["opacity"]),
                       torch.cat([prm["f_dc"], prm["f_rest"]], 1), cam)
        assert bool(pr["visible"].all())
        assert float((means[:, 0] / means[:, 2]).abs().max()) < 1.3 * cam.tanfovx
        assert float((means[:, 1] / means[:, 2]).abs().max()) < 1.3 * cam.tanfovy
        assert float(torch.sigmoid(prm["opacity"]).max()) < 0.99
        a, b, c = pr["cov2d"].unbind(-1)
        assert float((a * c - b * b).min()) > 0.25
    L = _loss(leaves, pose_l, cam, w)
    L.backward()
    assert float(L.detach()) > 1.0
    grads = {k: v.grad.clone() for k, v in leaves.items()}
    grads["pose"] = pose_l.grad.clone()
    g = torch.Generator().manual_seed(0)
    eps = 1e-6

    def fd(name, direction):
        vals = []
        for sgn in (+1.0, -1.0):
            p2 = {k: v.clone() for k, v in prm.items()}
            po2 = pose.clone()
            if name == "pose":
                po2 = po2 + sgn * eps * direction
            else:
                p2[name] = p2[name] + sgn * eps * direction
            with torch.no_grad():
                vals.append(float(_loss(p2, po2, cam, w)))
        return (vals[0] - vals[1]) / (2 * eps)

    worst = {}
    for name, gr in grads.items():
        scale = float(gr.abs().max())
        assert scale > 0, name
        errs = []
        # every pose coordinate; 24 seeded single coordinates of each per-Gaussian tensor (largest-gradient entry
        # included); 6 dense random directions
        if name == "pose":
            coords = list(range(7))
        else:
            coords = torch.randperm(gr.numel(), generator=g)[:23].tolist() + [int(gr.abs().reshape(-1).argmax())]
        for ci in coords:
            d = torch.zeros(gr.numel(), dtype=torch.float64)
            d[ci] = 1.0
            d = d.reshape(gr.shape)
            errs.append(abs(fd(name, d) - float((gr * d).sum())) / scale)
        for _ in range(6):
            d = torch.randn(gr.shape, generator=g, dtype=torch.float64)
            d = d / d.norm()
            ref = float((gr * d).sum())
            errs.append(abs(fd(name, d) - ref) / max(abs(ref), scale))
        worst[name] = max(errs)
    bad = {k: v for k, v in worst.items() if not v < TOL}
    assert not bad, f"oracle autograd vs fp64 central differences: {worst}"


def test_oracle_threshold_quirks_are_not_the_smooth_derivative():
    """Sanity of the test above: with the reference's cut-offs ON the forward changes (pairs below 1/255 are
    dropped), i.e. the smooth-regime switches really switch something."""
    prm, pose, cam, w = _scene()
    with torch.no_grad():
        a, _ = O.render_instantsplat(prm["xyz"], prm["rotation"], prm["scaling"], prm["opacity"], prm["f_dc"],
                                     prm["f_rest"], pose, cam, alpha_min=0.0, t_eps=0.0)
        b, _ = O.render_instantsplat(prm["xyz"], prm["rotation"], prm["scaling"], prm["opacity"], prm["f_dc"],
                                     prm["f_rest"], pose, cam)
    d = float((a - b).abs().max())
    assert 0.0 < d < 0.05
