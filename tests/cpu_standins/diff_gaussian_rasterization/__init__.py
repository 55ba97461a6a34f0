"""CPU stand-in for the `diff_gaussian_rasterization` package, TEST INFRASTRUCTURE ONLY: same public surface as
shims/diff_gaussian_rasterization (the NamedTuple is imported from the product so the field order is the product's),
with the rasterizer itself served by the CPU oracle.  Lets the reference's REAL Python (gaussian_renderer/__init__.py,
scene/gaussian_model.py, the train.py loop body) run here, where the reference tree exists but no GPU does."""
import torch
from torch import nn

from instantsplat_b200.rasterizer import GaussianRasterizationSettings  # noqa: F401
from oracle import gs_oracle as O


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        cam = O.Camera(int(rs.image_width), int(rs.image_height), float(rs.tanfovx), float(rs.tanfovy), rs.viewmatrix,
                       rs.projmatrix, rs.campos, rs.bg, int(rs.sh_degree), float(rs.scale_modifier))
        img, radii = O.rasterize(means3D, scales, rotations, opacities, shs, cam, means2D=means2D,
                                 colors_precomp=colors_precomp, cov3D_precomp=cov3D_precomp)
        return img, radii

    def markVisible(self, positions):
        with torch.no_grad():
            V = self.raster_settings.viewmatrix
            z = positions[:, 0] * V[0, 2] + positions[:, 1] * V[1, 2] + positions[:, 2] * V[2, 2] + V[3, 2]
            return z > 0.2
