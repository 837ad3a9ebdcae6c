"""Camera matrices of the reference (src/SwRast/Camera.h), in float32.

The 16 floats of ObjectToClip are an INPUT of the hot path: whoever drives the oracle and the
CUDA path computes them once here and hands both the same bits.
Matrices are column-major like glm: m[c, r] in a (4, 4) float32 array whose flat order is
glm's memory order, i.e. `m.reshape(16)[c * 4 + r]`.
"""
from __future__ import annotations

import math
import numpy as np

f32 = np.float32


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """glm-style mat4 * mat4 on column-major (4,4)[c,r] arrays, float32 left-to-right sums."""
    r = np.zeros((4, 4), dtype=f32)
    for c in range(4):
        for k in range(4):
            acc = f32(a[0, k]) * f32(b[c, 0])
            acc = f32(acc + f32(a[1, k]) * f32(b[c, 1]))
            acc = f32(acc + f32(a[2, k]) * f32(b[c, 2]))
            acc = f32(acc + f32(a[3, k]) * f32(b[c, 3]))
            r[c, k] = acc
    return r


def identity() -> np.ndarray:
    return np.eye(4, dtype=f32)


def translate(v) -> np.ndarray:
    m = identity()
    m[3, 0:3] = np.asarray(v, dtype=f32)
    return m


def scale(s) -> np.ndarray:
    m = identity()
    s = np.broadcast_to(np.asarray(s, dtype=f32), (3,))
    m[0, 0], m[1, 1], m[2, 2] = s
    return m


def rotate_axis(axis, angle) -> np.ndarray:
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    c, s = math.cos(angle), math.sin(angle)
    x, y, z = axis
    R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]])
    m = identity()
    m[0:3, 0:3] = R.T.astype(f32)  # column-major: m[c, r] = R[r, c]
    return m


class Camera:
    """Camera.h:7-48 — first-person camera with the reverse-Z infinite-far projection."""

    def __init__(self, position=(0.0, 0.0, 0.0), euler=(0.0, 0.0), fov_deg=90.0, aspect=1.0, near_z=0.01):
        self.position = np.asarray(position, dtype=np.float64)
        self.euler = (float(euler[0]), float(euler[1]))  # yaw, pitch
        self.fov = float(fov_deg)
        self.aspect = float(aspect)
        self.near_z = float(near_z)

    def rotation_quat(self):
        """Camera.h:58-63: rotateX(-pitch) * rotateY(yaw) as (w, x, y, z)."""
        yaw, pitch = self.euler
        sx, cx = math.sin(pitch * -0.5), math.cos(pitch * -0.5)
        sy, cy = math.sin(yaw * 0.5), math.cos(yaw * 0.5)
        return (cx * cy, sx * cy, cx * sy, sx * sy)

    def view_matrix(self, translate_to_origin=True) -> np.ndarray:
        """Camera.h:33-37: mat4_cast(ViewRotation), then translate(-ViewPosition)."""
        w, x, y, z = self.rotation_quat()
        m = identity()
        # glm::mat3_cast
        m[0, 0] = 1 - 2 * (y * y + z * z); m[0, 1] = 2 * (x * y + w * z); m[0, 2] = 2 * (x * z - w * y)
        m[1, 0] = 2 * (x * y - w * z); m[1, 1] = 1 - 2 * (x * x + z * z); m[1, 2] = 2 * (y * z + w * x)
        m[2, 0] = 2 * (x * z + w * y); m[2, 1] = 2 * (y * z - w * x); m[2, 2] = 1 - 2 * (x * x + y * y)
        if translate_to_origin:
            m = mat_mul(m, translate(-self.position))
        return m

    def proj_matrix(self) -> np.ndarray:
        """Camera.h:38-48."""
        f = 1.0 / math.tan(math.radians(self.fov) / 2.0)
        m = np.zeros((4, 4), dtype=f32)
        m[0, 0] = f / self.aspect
        m[1, 1] = -f
        m[2, 3] = -1.0
        m[3, 2] = self.near_z
        return m


def object_to_clip(proj: np.ndarray, view: np.ndarray, model: np.ndarray) -> np.ndarray:
    """ShadingContext::UpdateProj (Shading.h:35-39): (P * V) * M."""
    return mat_mul(mat_mul(proj, view), model)


def inverse_screen_proj(world_to_clip: np.ndarray, width: int, height: int, subpixel=(0.5, 0.5)) -> np.ndarray:
    """GetInverseScreenProjMatrix (Camera.h:140-146)."""
    inv = np.linalg.inv(world_to_clip.astype(np.float64).T).T.astype(f32)  # column-major storage
    inv = mat_mul(inv, translate((-1.0, -1.0, 0.0)))
    inv = mat_mul(inv, scale((2.0 / width, 2.0 / height, 1.0)))
    inv = mat_mul(inv, translate((subpixel[0], subpixel[1], 0.0)))
    return inv
