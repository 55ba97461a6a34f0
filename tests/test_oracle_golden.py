"""Pin the CPU oracle against vectors generated from the reference's own Python modules
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import torch

from oracle import gs_oracle as O


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_eval_sh_matches_reference(golden_dir):
    z = _g(golden_dir, "eval_sh.npz")
    sh, dirs = torch.from_numpy(z["sh"]), torch.from_numpy(z["dirs"])
    for d in range(4):
        out = O.eval_sh(d, sh, dirs)
        np.testing.assert_allclose(out.numpy(), z[f"deg{d}"], rtol=1e-6, atol=1e-6)


def test_loss_matches_reference(golden_dir):
    z = _g(golden_dir, "loss.npz")
    a = torch.from_numpy(z["img1"]).requires_grad_(True)
    b = torch.from_numpy(z["img2"])
    s = O.ssim(a, b)
    assert abs(s.item() - float(z["ssim"])) < 1e-6
    (gs,) = torch.autograd.grad(s, a)
    np.testing.assert_allclose(gs.numpy(), z["ssim_grad"], rtol=1e-4, atol=1e-8)
    l1 = O.l1_loss(a, b)
    assert abs(l1.item() - float(z["l1"])) < 1e-7
    tot = O.training_loss(a, b)
    assert abs(tot.item() - float(z["total"])) < 1e-6
    (gt,) = torch.autograd.grad(tot, a)
    np.testing.assert_allclose(gt.numpy(), z["total_grad"], rtol=1e-4, atol=1e-8)


def test_pose_helpers_match_reference(golden_dir):
    z = _g(golden_dir, "pose.npz")
    poses = torch.from_numpy(z["poses"])
    for i in range(poses.shape[0]):
        np.testing.assert_allclose(O.pose_to_w2c(poses[i]).numpy(), z["w2c"][i], rtol=1e-6, atol=1e-6)
    qm = O.quadmultiply(poses[0, :4], torch.from_numpy(z["q2"]))
    np.testing.assert_allclose(qm.numpy(), z["qmul"], rtol=1e-6, atol=1e-6)
    means, rots = O.pose_pretransform(torch.from_numpy(z["xyz"]), torch.from_numpy(z["q2"]), poses[0])
    np.testing.assert_allclose(means.numpy(), z["xyz_trans"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rots.numpy(), z["qmul"], rtol=1e-6, atol=1e-6)


def test_projection_matrix_matches_reference(golden_dir):
    z = _g(golden_dir, "proj.npz")
    P = O.projection_matrix(float(z["znear"]), float(z["zfar"]), float(z["fovx"]), float(z["fovy"]))
    np.testing.assert_allclose(P.numpy(), z["P"], rtol=1e-6, atol=1e-7)


def test_per_point_adam_matches_reference(golden_dir):
    z = _g(golden_dir, "per_point_adam.npz")
    a, b = torch.from_numpy(z["p_pp"]).clone(), torch.from_numpy(z["p_pl"]).clone()
    lr_pp = torch.from_numpy(z["lr_pp"])
    am, av = torch.zeros_like(a), torch.zeros_like(a)
    bm, bv = torch.zeros_like(b), torch.zeros_like(b)
    for it in range(6):
        O.per_point_adam_step(a, torch.from_numpy(z[f"ga{it}"]), am, av, it + 1, 1.6e-4,
                              per_point_lr=lr_pp)
        O.per_point_adam_step(b, torch.from_numpy(z[f"gb{it}"]), bm, bv, it + 1, 2.5e-2)
        for mine, ref in ((a, f"a{it}"), (am, f"am{it}"), (av, f"av{it}"), (b, f"b{it}"),
                          (bm, f"bm{it}"), (bv, f"bv{it}")):
            np.testing.assert_allclose(mine.numpy(), z[ref], rtol=1e-6, atol=1e-12)


def test_raster_regression_pin(golden_dir):
    """Oracle-vs-its-own committed fp64 output (NOT reference derived -- parity unpinned)."""
    z = _g(golden_dir, "raster_tiny.npz")
    W, H = int(z["width"]), int(z["height"])
    cam = O.Camera.instantsplat(W, H, float(z["fovx"]), float(z["fovy"]), bg=torch.from_numpy(z["bg"]),
                                sh_degree=3)
    P = {k: torch.from_numpy(z["in_" + k]) for k in ("xyz", "rotation", "scaling", "opacity", "f_dc", "f_rest")}
    img, radii = O.render_instantsplat(P["xyz"], P["rotation"], P["scaling"], P["opacity"], P["f_dc"],
                                       P["f_rest"], torch.from_numpy(z["pose"]), cam)
    bad = (img.numpy() - z["image"]).__abs__().max(0) > 1e-4
    assert (bad & ~z["ambiguous"]).sum() == 0
    assert (radii.numpy() != z["radii"]).mean() < 0.01
