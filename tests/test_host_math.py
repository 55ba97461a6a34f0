"""CPU check of the kernels' per-Gaussian arithmetic (instantsplat_b200/csrc/gs_math.cuh, compiled
for the host by g++) against the oracle: forward projection, analytic backward incl. the fused
pose gradient, and the lossless-culling predicate.  No GPU, no product library involved."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import gs_oracle as O
from instantsplat_b200.scenes import random_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class HostCam(ctypes.Structure):
    _fields_ = [("W", ctypes.c_int), ("H", ctypes.c_int), ("D", ctypes.c_int), ("M", ctypes.c_int),
                ("pose_on", ctypes.c_int), ("raw_params", ctypes.c_int),
                ("tanfovx", ctypes.c_float), ("tanfovy", ctypes.c_float), ("scale_mod", ctypes.c_float),
                ("V", ctypes.c_float * 16), ("Pm", ctypes.c_float * 16), ("campos", ctypes.c_float * 3),
                ("pose", ctypes.c_float * 7)]


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhostmath.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "host_math.cpp"), "-o", out])
    return ctypes.CDLL(out)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _cam(sc, deg, cam):
    h = HostCam()
    h.W, h.H, h.D, h.M, h.pose_on, h.raw_params = sc.width, sc.height, deg, 16, 1, 1
    h.tanfovx, h.tanfovy, h.scale_mod = cam.tanfovx, cam.tanfovy, 1.0
    h.V[:] = cam.viewmatrix.reshape(-1).tolist()
    h.Pm[:] = cam.projmatrix.reshape(-1).tolist()
    h.campos[:] = cam.campos.tolist()
    h.pose[:] = sc.poses[0].tolist()
    return h


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_forward_and_backward_match_oracle(lib, deg):
    sc = random_scene(1500, 96, 80, seed=11 + deg, sh_degree=deg)
    cam = O.Camera.instantsplat(sc.width, sc.height, sc.fovx, sc.fovy, sh_degree=deg)
    P = sc.P
    prm = {k: v.clone().requires_grad_(True) for k, v in sc.params.items()}
    pose = sc.poses[0].clone().requires_grad_(True)
    m2d = torch.zeros(P, 3, requires_grad=True)
    means, rots = O.pose_pretransform(prm["xyz"], prm["rotation"], pose)
    shs = torch.cat([prm["f_dc"], prm["f_rest"]], dim=1)
    proj = O.project(means, torch.exp(prm["scaling"]), rots, torch.sigmoid(prm["opacity"]), shs, cam,
                     means2D=m2d)
    vis = proj["visible"]

    h = _cam(sc, deg, cam)
    np_in = {k: np.ascontiguousarray(v.numpy()) for k, v in sc.params.items()}
    geom = np.zeros((P, 12), np.float32)
    rect = np.zeros((P, 4), np.int32)
    clamped = np.zeros(P, np.int32)
    lib.host_project(ctypes.byref(h), P, _fp(np_in["xyz"]), _fp(np_in["scaling"]), _fp(np_in["rotation"]),
                     _fp(np_in["opacity"]), _fp(np_in["f_dc"]), _fp(np_in["f_rest"]), _fp(geom), _ip(rect),
                     _ip(clamped))
    v = vis.numpy()
    assert (geom[:, 11] > 0).sum() > 200
    # visibility / radii / rect: identical except for fp-rounding borderline cases
    assert ((geom[:, 11] > 0) != v).mean() < 2e-3
    both = v & (geom[:, 11] > 0)
    assert (geom[both, 10] != proj["radii"].numpy()[both]).mean() < 2e-3
    assert (rect[both] != proj["rect"].numpy()[both]).any(1).mean() < 5e-3
    ref = torch.cat([proj["xy"], proj["conic"], proj["opacity"][:, None], proj["depth"][:, None],
                     proj["rgb"]], dim=1).detach().numpy()
    np.testing.assert_allclose(geom[both, :10], ref[both], rtol=2e-4, atol=2e-4)

    # ---- backward with random cotangents on the 9 splat quantities
    g = torch.Generator().manual_seed(3)
    ds = torch.randn(P, 9, generator=g) * vis[:, None]
    L = (ds[:, 0:2] * proj["xy"]).sum() + (ds[:, 2:5] * proj["conic"]).sum() \
        + (ds[:, 5] * proj["opacity"]).sum() + (ds[:, 6:9] * proj["rgb"]).sum()
    L.backward()
    dsn = np.ascontiguousarray(ds.numpy())
    dm = np.zeros((P, 3), np.float32); dsc = np.zeros((P, 3), np.float32); dq = np.zeros((P, 4), np.float32)
    dop = np.zeros(P, np.float32); ddc = np.zeros((P, 3), np.float32); drest = np.zeros((P, 15, 3), np.float32)
    dm2 = np.zeros((P, 2), np.float32); dpose = np.zeros(7, np.float32)
    lib.host_project_bwd(ctypes.byref(h), P, _fp(np_in["xyz"]), _fp(np_in["scaling"]), _fp(np_in["rotation"]),
                         _fp(np_in["opacity"]), _fp(np_in["f_dc"]), _fp(np_in["f_rest"]), _fp(dsn), _fp(dm),
                         _fp(dsc), _fp(dq), _fp(dop), _fp(ddc), _fp(drest), _fp(dm2), _fp(dpose))

    def close(mine, ref_t, name, rtol=2e-3):
        r = ref_t.numpy().reshape(mine.shape)
        sel = both if mine.shape[0] == P else slice(None)
        scale = np.abs(r[sel]).max() + 1e-12
        err = np.abs(mine[sel] - r[sel]).max() / scale
        assert err < rtol, (name, err)

    close(dm, prm["xyz"].grad, "xyz")
    close(dsc, prm["scaling"].grad, "scaling")
    close(dq, prm["rotation"].grad, "rotation")
    close(dop, prm["opacity"].grad, "opacity")
    close(ddc, prm["f_dc"].grad, "f_dc")
    close(drest, prm["f_rest"].grad, "f_rest")
    close(dm2, m2d.grad[:, :2], "means2D")
    close(dpose, pose.grad, "pose")


def test_cull_predicate_is_conservative(lib):
    """rect_may_contribute must never reject a rectangle containing a pixel with alpha >= 1/255."""
    rng = np.random.default_rng(0)
    lib.host_rect_may_contribute.argtypes = [ctypes.c_float] * 10
    n_rej = 0
    for _ in range(4000):
        a, c = rng.uniform(0.3, 40), rng.uniform(0.3, 40)
        b = rng.uniform(-0.95, 0.95) * math.sqrt(a * c)
        det = a * c - b * b
        A, B, C = c / det, -b / det, a / det
        op = rng.uniform(0.0, 1.0) ** 2
        x, y = rng.uniform(-20, 40), rng.uniform(-20, 40)
        w, hgt = rng.integers(1, 17), rng.integers(1, 17)
        rx0, ry0 = rng.integers(0, 16), rng.integers(0, 16)
        xs, ys = np.meshgrid(np.arange(rx0, rx0 + w), np.arange(ry0, ry0 + hgt))
        dx, dy = x - xs, y - ys
        power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
        alpha = np.minimum(0.99, op * np.exp(power))
        contributes = bool(((power <= 0) & (alpha >= 1 / 255)).any())
        keep = lib.host_rect_may_contribute(x, y, A, B, C, op, rx0, ry0, rx0 + w - 1, ry0 + hgt - 1)
        if contributes:
            assert keep == 1
        n_rej += (keep == 0)
    assert n_rej > 500          # the predicate does cull


def test_row_cull_is_conservative_and_agrees_with_the_tile_predicate(lib):
    """row_keep_range (one interval per tile row; what k_preprocess uses for rects of <= 24 tiles) against (a) brute
    force over the pixels -- a tile holding a pixel with alpha >= 1/255 must be kept -- and (b) the per-tile predicate
    rect_may_contribute, which is the same set in exact arithmetic (disagreements only on borderline tiles)."""
    rng = np.random.default_rng(1)
    lib.host_rect_may_contribute.argtypes = [ctypes.c_float] * 10
    lib.host_row_keep.argtypes = [ctypes.c_float] * 6 + [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_int)]
    W, H = 30 * 16 - 5, 20 * 16 - 9              # last tile row / column are partial
    gx, gy = 30, 20
    n_tiles = n_kept = n_diff = n_must = 0
    for it in range(3000):
        s1, s2 = rng.uniform(0.3, 30), rng.uniform(0.3, 30)
        if it % 5 == 0:
            s2 = s1 * rng.uniform(0.02, 0.2)             # needle-shaped
        th = rng.uniform(0, math.pi)
        a = (math.cos(th) * s1) ** 2 + (math.sin(th) * s2) ** 2 + 0.3
        c = (math.sin(th) * s1) ** 2 + (math.cos(th) * s2) ** 2 + 0.3
        b = math.cos(th) * math.sin(th) * (s1 * s1 - s2 * s2)
        det = a * c - b * b
        A, B, C = c / det, -b / det, a / det
        op = rng.uniform(0.004, 1.0) if it % 7 else rng.uniform(0.0, 0.006)
        x, y = rng.uniform(-30, W + 30), rng.uniform(-30, H + 30)
        rad = math.ceil(3 * math.sqrt(0.5 * (a + c) + math.sqrt(max(0.1, (0.5 * (a + c)) ** 2 - det))))
        rx0, rx1 = min(gx, max(0, int((x - rad) / 16))), min(gx, max(0, int((x + rad + 15) / 16)))
        ry0, ry1 = min(gy, max(0, int((y - rad) / 16))), min(gy, max(0, int((y + rad + 15) / 16)))
        w, h = rx1 - rx0, ry1 - ry0
        if w * h == 0:
            continue
        kept = (ctypes.c_int * (w * h))()
        ok = lib.host_row_keep(x, y, A, B, C, op, H, rx0, ry0, rx1, ry1, kept)
        if not ok:
            assert op * 255 < 0.999          # only "can never reach 1/255" is allowed to be degenerate here
            continue
        for ty in range(ry0, ry1):
            for tx in range(rx0, rx1):
                x0, y0 = tx * 16, ty * 16
                x1, y1 = min(x0 + 15, W - 1), min(y0 + 15, H - 1)
                xs, ys = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1))
                dx, dy = x - xs, y - ys
                power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
                must = bool(((power <= 0) & (op * np.exp(power) >= 1 / 255)).any())
                k = kept[(ty - ry0) * w + (tx - rx0)]
                if must:
                    assert k == 1, (it, tx, ty, x, y, A, B, C, op)
                t = lib.host_rect_may_contribute(x, y, A, B, C, op, x0, y0, x1, y1)
                n_tiles += 1
                n_kept += k
                n_must += must
                n_diff += int(k != t)
    assert n_tiles > 20000 and n_must > 3000
    assert n_kept < 0.9 * n_tiles                 # it does cull
    assert n_diff <= 0.003 * n_tiles, (n_diff, n_tiles)
