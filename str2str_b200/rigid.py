"""Minimal `Rigid` / `Rotation` value types that honour the reference's call signatures at the boundary.

The reference passes frames around as `Rigid` objects (src/common/rigid_utils.py:856-1451).  Inside this
package frames live as tensor_7 ([quat wxyz, trans]) in HBM; these classes only wrap such tensors so that
`out['rigids']`, `diffuser.score(rigids_0=..., rigids_t=...)` and `diffuser.reverse(...)` keep their reference
shapes.  Conversions are a few elementwise torch ops on the tensor's own device (boundary glue, not the hot path).
"""
from __future__ import annotations

from typing import Optional

import torch


def _quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    # quadratic form without normalisation (rigid_utils.py:187-207)
    a, b, c, d = q.unbind(-1)
    rows = [
        a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
        2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
        2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d,
    ]
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def _rot_to_quat(m: torch.Tensor) -> torch.Tensor:
    # rotation3d.matrix_to_quaternion :102-161 (what Rotation.get_quats uses, rigid_utils.py:541)
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]
    q_abs = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1).clamp(min=0).sqrt()
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    ) / (2.0 * q_abs[..., None].clamp(min=0.1))
    pick = q_abs.argmax(-1)
    return torch.gather(cand, -2, pick[..., None, None].expand(pick.shape + (1, 4))).squeeze(-2)


class Rotation:
    def __init__(self, rot_mats: Optional[torch.Tensor] = None, quats: Optional[torch.Tensor] = None, normalize_quats: bool = True):
        if (rot_mats is None) == (quats is None):
            raise ValueError("Exactly one input argument must be specified")
        if (rot_mats is not None and rot_mats.shape[-2:] != (3, 3)) or (quats is not None and quats.shape[-1] != 4):
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        if quats is not None:
            quats = quats.to(torch.float32)
            if normalize_quats:
                quats = quats / torch.linalg.norm(quats, dim=-1, keepdim=True)
        else:
            rot_mats = rot_mats.to(torch.float32)
        self._rot_mats, self._quats = rot_mats, quats

    @property
    def shape(self):
        return self._rot_mats.shape[:-2] if self._rot_mats is not None else self._quats.shape[:-1]

    @property
    def device(self):
        return (self._rot_mats if self._rot_mats is not None else self._quats).device

    def get_rot_mats(self) -> torch.Tensor:
        return self._rot_mats if self._rot_mats is not None else _quat_to_rot(self._quats)

    def get_quats(self) -> torch.Tensor:
        return self._quats if self._quats is not None else _rot_to_quat(self._rot_mats)


class Rigid:
    def __init__(self, rots: Optional[Rotation], trans: Optional[torch.Tensor]):
        if rots is None and trans is None:
            raise ValueError("At least one input argument must be specified")
        if rots is None:
            eye = torch.eye(3, device=trans.device).expand(*trans.shape[:-1], 3, 3)
            rots = Rotation(rot_mats=eye)
        if trans is None:
            trans = torch.zeros(*rots.shape, 3, device=rots.device)
        if rots.shape != trans.shape[:-1] or rots.device != trans.device:
            raise ValueError("Rots and trans incompatible")
        self._rots, self._trans = rots, trans.to(torch.float32)

    @property
    def shape(self):
        return self._trans.shape[:-1]

    @property
    def device(self):
        return self._trans.device

    def get_rots(self) -> Rotation:
        return self._rots

    def get_trans(self) -> torch.Tensor:
        return self._trans

    def to_tensor_7(self) -> torch.Tensor:
        return torch.cat([self._rots.get_quats(), self._trans], -1)

    @staticmethod
    def from_tensor_7(t: torch.Tensor, normalize_quats: bool = False) -> "Rigid":
        if t.shape[-1] != 7:
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(quats=t[..., :4], normalize_quats=normalize_quats), t[..., 4:])

    @staticmethod
    def from_tensor_4x4(t: torch.Tensor) -> "Rigid":
        if t.shape[-2:] != (4, 4):
            raise ValueError("Incorrectly shaped input tensor")
        return Rigid(Rotation(rot_mats=t[..., :3, :3]), t[..., :3, 3])

    def apply_trans_fn(self, fn) -> "Rigid":
        return Rigid(self._rots, fn(self._trans))
