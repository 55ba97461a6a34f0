"""Minimal stand-in for the attributes `render()` and the optimizer set-up read from the reference's
`GaussianModel` (/root/reference/scene/gaussian_model.py:101-136,203-221), so the drop-in boundary
(render -> loss -> backward -> PerPointAdam.step, i.e. the body of /root/reference/train.py:140-211) can be
exercised and timed without the reference's file-based scene loading."""
from __future__ import annotations

import torch

from .per_point_adam import PerPointAdam
from .scenes import Scene


class SimpleGaussianModel:
    def __init__(self, scene: Scene, device="cuda", sh_degree=None):
        p = {k: torch.nn.Parameter(v.to(device).float().contiguous()) for k, v in scene.params.items()}
        self._xyz, self._rotation, self._scaling, self._opacity = p["xyz"], p["rotation"], p["scaling"], p["opacity"]
        self._features_dc, self._features_rest = p["f_dc"], p["f_rest"]
        self.max_sh_degree = 3
        self.active_sh_degree = scene.sh_degree if sh_degree is None else sh_degree
        self.P = torch.nn.Parameter(scene.poses.to(device).float().contiguous())
        self.per_point_lr = None if scene.per_point_lr is None else scene.per_point_lr.to(device).float().reshape(-1, 1)
        self.optimizer = None

    get_xyz = property(lambda s: s._xyz)
    get_opacity = property(lambda s: torch.sigmoid(s._opacity))
    get_scaling = property(lambda s: torch.exp(s._scaling))
    get_rotation = property(lambda s: torch.nn.functional.normalize(s._rotation))
    get_features = property(lambda s: torch.cat((s._features_dc, s._features_rest), dim=1))

    def get_RT(self, idx):
        return self.P[idx]

    def training_setup_pp(self, position_lr=0.00016, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005,
                          rotation_lr=0.001):
        """Param groups of /root/reference/scene/gaussian_model.py:203-221."""
        groups = [
            {"params": [self._xyz], "per_point_lr": self.per_point_lr, "lr": position_lr, "name": "xyz"},
            {"params": [self._features_dc], "lr": feature_lr * 10, "name": "f_dc"},
            {"params": [self._features_rest], "lr": feature_lr / 20.0 * 10, "name": "f_rest"},
            {"params": [self._opacity], "lr": opacity_lr, "name": "opacity"},
            {"params": [self._scaling], "lr": scaling_lr * 10, "name": "scaling"},
            {"params": [self._rotation], "lr": rotation_lr * 10, "name": "rotation"},
            {"params": [self.P], "lr": rotation_lr * 0.1, "name": "pose"},
        ]
        self.optimizer = PerPointAdam(groups, lr=0, betas=(0.9, 0.999), eps=1e-15, weight_decay=0.0)
        return self.optimizer


class PipelineDefaults:
    """/root/reference/arguments/__init__.py:66-71"""
    convert_SHs_python = False
    compute_cov3D_python = False
    debug = False
