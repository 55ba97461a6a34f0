"""On-disk formats either side of the hot path -- host-side, numpy only.

What the reference's training stage reads and writes, so a scene produced here can drive an unchanged
`train.py` and its outputs can be loaded back:

* `sparse_<n>/0/cameras.txt` (PINHOLE only) and `images.txt` (two lines per image), as parsed by
  /root/reference/scene/colmap_loader.py:159-182,248-275 and written by
  /root/reference/utils/sfm_utils.py:202-315;
* `points3D.ply`: x,y,z f4 | nx,ny,nz f4 | red,green,blue u1 (/root/reference/scene/dataset_readers.py:315-369);
* `confidence_dsp.npy` [P,1] (/root/reference/train.py:63-96);
* the trained model `point_cloud.ply`: x,y,z,nx,ny,nz,f_dc_0..2,f_rest_0..44,opacity,scale_0..2,rot_0..3, all f4,
  SH stored channel-major (/root/reference/scene/gaussian_model.py:247-278);
* `pose_optimized.npy` [n,4,4] world-to-camera matrices ordered by COLMAP image id
  (/root/reference/train.py:46-60).

PLY files are written/read directly (binary little endian) -- `plyfile` is not a dependency.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------
# quaternion <-> rotation (COLMAP convention: qvec = (w, x, y, z) of the world-to-camera rotation)
# ----------------------------------------------------------------------------------------------
def qvec2rotmat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def pose7_to_matrix(pose: np.ndarray) -> np.ndarray:
    """[qw,qx,qy,qz,tx,ty,tz] -> 4x4 world-to-camera; the quaternion is normalised first, like
    get_camera_from_tensor (/root/reference/utils/pose_utils.py:57-84)."""
    q = np.asarray(pose[:4], dtype=np.float64)
    q = q / np.linalg.norm(q)
    M = np.eye(4)
    M[:3, :3] = qvec2rotmat(q)
    M[:3, 3] = pose[4:7]
    return M


# ----------------------------------------------------------------------------------------------
# COLMAP text model
# ----------------------------------------------------------------------------------------------
def write_colmap_text(folder: str, width: int, height: int, fx: float, fy: float, poses: np.ndarray,
                      image_names: List[str], cx: float = None, cy: float = None) -> None:
    """One shared PINHOLE camera, one image per pose row [qw,qx,qy,qz,tx,ty,tz] (world-to-camera)."""
    os.makedirs(folder, exist_ok=True)
    cx = width / 2.0 if cx is None else cx
    cy = height / 2.0 if cy is None else cy
    with open(os.path.join(folder, "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        f.write(f"# Number of cameras: {len(image_names)}\n")
        for i in range(len(image_names)):
            f.write(f"{i + 1} PINHOLE {width} {height} {fx:.10g} {fy:.10g} {cx:.10g} {cy:.10g}\n")
    with open(os.path.join(folder, "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n"
                "#   POINTS2D[] as (X, Y, POINT3D_ID)\n")
        for i, (p, name) in enumerate(zip(poses, image_names)):
            q = np.asarray(p[:4], dtype=np.float64)
            q = q / np.linalg.norm(q)
            vals = " ".join(f"{v:.10g}" for v in list(q) + list(p[4:7]))
            f.write(f"{i + 1} {vals} {i + 1} {name}\n\n")


def read_colmap_text(folder: str) -> Tuple[Dict[int, dict], Dict[int, dict]]:
    """Returns (cameras, images) keyed by id; same fields the reference's readers expose."""
    cameras, images = {}, {}
    with open(os.path.join(folder, "cameras.txt")) as f:
        for line in f:
            line = line.strip()
            if not line or line[0] == "#":
                continue
            e = line.split()
            if e[1] != "PINHOLE":
                raise ValueError("only PINHOLE cameras are supported (as in the reference)")
            cameras[int(e[0])] = dict(id=int(e[0]), model=e[1], width=int(e[2]), height=int(e[3]),
                                      params=np.array(list(map(float, e[4:]))))
    with open(os.path.join(folder, "images.txt")) as f:
        lines = f.readlines()
    i = 0
    while i < len(lines):
        line = lines[i].strip()
        i += 1
        if not line or line[0] == "#":
            continue
        e = line.split()
        images[int(e[0])] = dict(id=int(e[0]), qvec=np.array(list(map(float, e[1:5]))),
                                 tvec=np.array(list(map(float, e[5:8]))), camera_id=int(e[8]), name=e[9])
        i += 1                                     # the POINTS2D line
    return cameras, images


# ----------------------------------------------------------------------------------------------
# PLY (binary little endian, single 'vertex' element)
# ----------------------------------------------------------------------------------------------
_PLY_TYPES = {"f4": "float", "u1": "uchar", "f8": "double", "i4": "int"}
_PLY_REV = {"float": "f4", "float32": "f4", "uchar": "u1", "uint8": "u1", "double": "f8", "float64": "f8",
            "int": "i4", "int32": "i4"}


def _write_ply(path: str, fields: List[Tuple[str, str]], columns: List[np.ndarray]) -> None:
    n = columns[0].shape[0]
    dt = np.dtype([(name, "<" + t) for name, t in fields])
    arr = np.empty(n, dtype=dt)
    for (name, _), col in zip(fields, columns):
        arr[name] = col
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {n}"]
        hdr += [f"property {_PLY_TYPES[t]} {name}" for name, t in fields]
        hdr.append("end_header")
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(arr.tobytes())


def _read_ply(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError("not a PLY file")
        fmt = f.readline().split()
        if fmt[1] != b"binary_little_endian":
            raise ValueError("only binary_little_endian PLY is supported")
        n, fields = None, []
        while True:
            line = f.readline().strip()
            if line == b"end_header":
                break
            tok = line.split()
            if tok[:2] == [b"element", b"vertex"]:
                n = int(tok[2])
            elif tok[0] == b"property":
                fields.append((tok[2].decode(), "<" + _PLY_REV[tok[1].decode()]))
        data = np.frombuffer(f.read(n * np.dtype(fields).itemsize), dtype=np.dtype(fields), count=n)
    return data


def write_points3d_ply(path: str, xyz: np.ndarray, rgb_u8: np.ndarray, normals: np.ndarray = None) -> None:
    normals = np.zeros_like(xyz) if normals is None else normals
    fields = [("x", "f4"), ("y", "f4"), ("z", "f4"), ("nx", "f4"), ("ny", "f4"), ("nz", "f4"),
              ("red", "u1"), ("green", "u1"), ("blue", "u1")]
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2], normals[:, 0], normals[:, 1], normals[:, 2],
            rgb_u8[:, 0], rgb_u8[:, 1], rgb_u8[:, 2]]
    _write_ply(path, fields, cols)


def read_points3d_ply(path: str):
    d = _read_ply(path)
    xyz = np.stack([d["x"], d["y"], d["z"]], axis=1)
    rgb = np.stack([d["red"], d["green"], d["blue"]], axis=1)
    nrm = np.stack([d["nx"], d["ny"], d["nz"]], axis=1) if "nx" in d.dtype.names else np.zeros_like(xyz)
    return xyz, rgb, nrm


def gaussian_ply_fields(n_rest: int = 45) -> List[str]:
    """construct_list_of_attributes (/root/reference/scene/gaussian_model.py:247-259)."""
    return (["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(n_rest)]
            + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)])


def save_gaussians_ply(path: str, xyz, f_dc, f_rest, opacity, scaling, rotation) -> None:
    """Model tensors -> point_cloud.ply; f_dc [P,1,3], f_rest [P,M-1,3] are stored channel-major
    (transpose(1,2).flatten(1)), raw (un-activated) opacity / scaling / rotation, zero normals."""
    xyz = np.asarray(xyz, np.float32)
    P = xyz.shape[0]
    dc = np.asarray(f_dc, np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    rest = np.asarray(f_rest, np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    attrs = np.concatenate([xyz, np.zeros_like(xyz), dc, rest, np.asarray(opacity, np.float32).reshape(P, 1),
                            np.asarray(scaling, np.float32).reshape(P, 3), np.asarray(rotation, np.float32).reshape(P, 4)],
                           axis=1)
    names = gaussian_ply_fields(rest.shape[1])
    _write_ply(path, [(n, "f4") for n in names], [attrs[:, i] for i in range(attrs.shape[1])])


def load_gaussians_ply(path: str) -> Dict[str, np.ndarray]:
    """Inverse of save_gaussians_ply (what GaussianModel.load_ply does, gaussian_model.py:280+)."""
    d = _read_ply(path)
    P = d.shape[0]
    col = lambda prefix: np.stack([d[n] for n in sorted((n for n in d.dtype.names if n.startswith(prefix)),
                                                        key=lambda s: int(s.split("_")[-1]))], axis=1)
    rest = col("f_rest_")
    return dict(xyz=np.stack([d["x"], d["y"], d["z"]], axis=1),
                f_dc=col("f_dc_").reshape(P, 3, 1).transpose(0, 2, 1).copy(),
                f_rest=rest.reshape(P, 3, rest.shape[1] // 3).transpose(0, 2, 1).copy(),
                opacity=d["opacity"].reshape(P, 1).copy(), scaling=col("scale_"), rotation=col("rot_"))


def save_pose_npy(path: str, poses7: np.ndarray, colmap_ids: List[int] = None) -> np.ndarray:
    """train.py:46-60: [n,7] quaternion poses -> [n,4,4] w2c matrices, row i = the camera with COLMAP id i+1."""
    n = poses7.shape[0]
    ids = list(range(1, n + 1)) if colmap_ids is None else list(colmap_ids)
    mats = np.stack([pose7_to_matrix(poses7[ids.index(i + 1)]) for i in range(n)]).astype(np.float32)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.save(path, mats)
    return mats


# ----------------------------------------------------------------------------------------------
# a complete synthetic `source_path` for an unchanged train.py
# ----------------------------------------------------------------------------------------------
def write_synthetic_source(root: str, scene, images_u8: np.ndarray = None) -> str:
    """Writes <root>/sparse_<n>/0/{cameras.txt,images.txt,points3D.ply,confidence_dsp.npy} (+ <root>/images/*.png when
    `images_u8` [n,H,W,3] is given and PIL is importable) for a `instantsplat_b200.scenes.Scene`."""
    import math
    n = scene.n_views
    folder = os.path.join(root, f"sparse_{n}", "0")
    fx = scene.width / (2.0 * math.tan(scene.fovx / 2.0))
    fy = scene.height / (2.0 * math.tan(scene.fovy / 2.0))
    names = [f"{i:06d}.png" for i in range(n)]
    write_colmap_text(folder, scene.width, scene.height, fx, fy, scene.poses.numpy(), names)
    xyz = scene.params["xyz"].numpy()
    rgb = np.clip((scene.params["f_dc"].numpy().reshape(-1, 3) * 0.28209479177387814 + 0.5) * 255.0, 0, 255).astype(np.uint8)
    write_points3d_ply(os.path.join(folder, "points3D.ply"), xyz, rgb)
    conf = np.zeros((scene.P, 1), np.float32)
    if scene.per_point_lr is not None:       # invert train.py:63-85: lr = (1 - sigmoid(c)) * 99 + 1
        s = np.clip((scene.per_point_lr.numpy().reshape(-1, 1) - 1.0) / 99.0, 1e-6, 1 - 1e-6)
        conf = np.log((1 - s) / s).astype(np.float32)
    np.save(os.path.join(folder, "confidence_dsp.npy"), conf)
    if images_u8 is not None:
        try:
            from PIL import Image
            os.makedirs(os.path.join(root, "images"), exist_ok=True)
            for i, name in enumerate(names):
                Image.fromarray(images_u8[i]).save(os.path.join(root, "images", name))
        except ImportError:
            pass
    return folder
