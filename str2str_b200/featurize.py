"""The step right BEFORE the hot path: PDB text -> the feature dict `predict_step` consumes.

Mirrors, for the keys the sampling path reads, the reference's
  * `protein.from_pdb_string`                       src/common/protein.py:72-140   (here: a column parser, no Biopython)
  * `ProteinFeatureTransform`                       src/data/components/dataset.py:26-143
  * `data_transforms.atom37_to_frames`              src/common/data_transforms.py:758-894
  * `data_transforms.atom37_to_torsion_angles`      src/common/data_transforms.py:925-1090
  * `data_transforms.get_backbone_frames`           src/common/data_transforms.py:1093-1100
  * `SamplingPDBDataset` / `BatchTensorConverter`   dataset.py:304-320, protein_datamodule.py:9-57

The arithmetic is restated as a handful of batched tensor expressions over all residues and all 8 rigid groups / 7
torsions at once (the reference walks `Rigid` objects and `batched_gather`); the rounding points that shape the result
are kept: frames are built in fp64 from the parsed coordinates, ROUNDED TO fp32 (the reference's `Rigid` forces fp32,
rigid_utils.py:327-331,902), and the torsion frames are inverted in fp32 before they meet the fp64 fourth atom.
One-off per protein, host side, torch CPU; pinned to the unmodified reference by tests/golden/featurize_*.npz.

All 34 tensors of the reference transform are produced (the sampling path reads seven of them; the chi / pseudo-beta / atom14
features are the training losses' inputs: data_transforms.py:370-402 `make_pseudo_beta`, :575-646 `make_atom14_masks`, :656-757
`make_atom14_positions`, :1103-1110 `get_chi_angles`).
"""
from __future__ import annotations

import glob
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

# --- residue tables (chemistry; residue_constants.py:34-111, 369-374, 492-537) ---------------------------------------
RESTYPES = "ARNDCQEGHILKMFPSTWYV"
RESNAMES = ("ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL").split()
RESNAME_TO_IDX = {n: i for i, n in enumerate(RESNAMES)}
ATOM_TYPES = ("N CA C CB O CG CG1 CG2 OG OG1 SG CD CD1 CD2 ND1 ND2 OD1 OD2 SD CE CE1 CE2 CE3 NE NE1 NE2 OE1 OE2 CH2 "
              "NH1 NH2 OH CZ CZ2 CZ3 NZ OXT").split()
ATOM_ORDER = {a: i for i, a in enumerate(ATOM_TYPES)}
CA_IDX = ATOM_ORDER["CA"]
# side-chain dihedral chains: chi_k is the dihedral of atoms k..k+3 of the chain
_CHI_CHAIN = {
    "ARG": "N CA CB CG CD NE CZ", "ASN": "N CA CB CG OD1", "ASP": "N CA CB CG OD1", "CYS": "N CA CB SG",
    "GLN": "N CA CB CG CD OE1", "GLU": "N CA CB CG CD OE1", "HIS": "N CA CB CG ND1", "ILE": "N CA CB CG1 CD1",
    "LEU": "N CA CB CG CD1", "LYS": "N CA CB CG CD CE NZ", "MET": "N CA CB CG SD CE", "PHE": "N CA CB CG CD1",
    "PRO": "N CA CB CG CD", "SER": "N CA CB OG", "THR": "N CA CB OG1", "TRP": "N CA CB CG CD1", "TYR": "N CA CB CG CD1",
    "VAL": "N CA CB CG1",
}
_PI_PERIODIC = {"ASP": 1, "GLU": 2, "PHE": 1, "TYR": 1}      # chi index whose 180-degree flip is a relabelling


def _tables():
    chi_idx = np.zeros((21, 4, 4), np.int64)                  # get_chi_atom_indices(): zeros where a chi is absent
    chi_mask = np.zeros((21, 4), np.float64)
    for r, name in enumerate(RESNAMES):
        chain = _CHI_CHAIN.get(name, "").split()
        for k in range(max(0, len(chain) - 3)):
            chi_idx[r, k] = [ATOM_ORDER[a] for a in chain[k:k + 4]]
            chi_mask[r, k] = 1.0
    pi = np.zeros((21, 4), np.float64)
    for name, k in _PI_PERIODIC.items():
        pi[RESNAME_TO_IDX[name], k] = 1.0
    # rigid groups: 0 backbone (C, CA, N), 3 psi (CA, C, O), 4.. chi groups (last three atoms of the dihedral)
    base = np.zeros((21, 8, 3), np.int64)
    base[:, 0] = [ATOM_ORDER["C"], ATOM_ORDER["CA"], ATOM_ORDER["N"]]
    base[:, 3] = [ATOM_ORDER["CA"], ATOM_ORDER["C"], ATOM_ORDER["O"]]
    base[:20, 4:] = np.where(chi_mask[:20, :, None] > 0, chi_idx[:20, :, 1:], 0)
    group_mask = np.zeros((21, 8), np.float64)
    group_mask[:, 0] = group_mask[:, 3] = 1.0
    group_mask[:20, 4:] = chi_mask[:20]
    ambiguous = np.zeros((21, 8), np.float64)
    for name in _PI_PERIODIC:                                 # residue_atom_renaming_swaps: the LAST chi group flips
        r = RESNAME_TO_IDX[name]
        ambiguous[r, int(chi_mask[r].sum()) - 1 + 4] = 1.0
    return chi_idx, chi_mask, pi, base, group_mask, ambiguous


CHI_ATOM_IDX, CHI_MASK, CHI_PI_PERIODIC, GROUP_BASE_ATOMS, GROUP_MASK, GROUP_AMBIGUOUS = _tables()

# heavy atoms of each residue type in atom14 order (residue_constants.py:504-527) and the pairs whose names are interchangeable (:369-374)
_ATOM14 = {
    "ALA": "N CA C O CB", "ARG": "N CA C O CB CG CD NE CZ NH1 NH2", "ASN": "N CA C O CB CG OD1 ND2", "ASP": "N CA C O CB CG OD1 OD2",
    "CYS": "N CA C O CB SG", "GLN": "N CA C O CB CG CD OE1 NE2", "GLU": "N CA C O CB CG CD OE1 OE2", "GLY": "N CA C O",
    "HIS": "N CA C O CB CG ND1 CD2 CE1 NE2", "ILE": "N CA C O CB CG1 CG2 CD1", "LEU": "N CA C O CB CG CD1 CD2",
    "LYS": "N CA C O CB CG CD CE NZ", "MET": "N CA C O CB CG SD CE", "PHE": "N CA C O CB CG CD1 CD2 CE1 CE2 CZ", "PRO": "N CA C O CB CG CD",
    "SER": "N CA C O CB OG", "THR": "N CA C O CB OG1 CG2", "TRP": "N CA C O CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2",
    "TYR": "N CA C O CB CG CD1 CD2 CE1 CE2 CZ OH", "VAL": "N CA C O CB CG1 CG2",
}
_SWAPS = {"ASP": {"OD1": "OD2"}, "GLU": {"OE1": "OE2"}, "PHE": {"CD1": "CD2", "CE1": "CE2"}, "TYR": {"CD1": "CD2", "CE1": "CE2"}}


def _atom14_tables():
    a14_to_37 = np.zeros((21, 14), np.int64)
    a37_to_14 = np.zeros((21, 37), np.int64)
    a14_mask = np.zeros((21, 14), np.float32)
    a37_mask = np.zeros((21, 37), np.float32)
    rename = np.tile(np.eye(14), (21, 1, 1))
    ambiguous = np.zeros((21, 14), np.float64)
    for r, name in enumerate(RESNAMES):
        names = _ATOM14[name].split()
        for i, a in enumerate(names):
            a14_to_37[r, i] = ATOM_ORDER[a]
            a37_to_14[r, ATOM_ORDER[a]] = i
            a14_mask[r, i] = 1.0
            a37_mask[r, ATOM_ORDER[a]] = 1.0
        for a, b in _SWAPS.get(name, {}).items():
            i, j = names.index(a), names.index(b)
            rename[r, i, i] = rename[r, j, j] = 0.0
            rename[r, i, j] = rename[r, j, i] = 1.0
            ambiguous[r, i] = ambiguous[r, j] = 1.0
    return a14_to_37, a37_to_14, a14_mask, a37_mask, rename, ambiguous


ATOM14_TO_37, ATOM37_TO_14, ATOM14_MASK, ATOM37_MASK, ATOM14_RENAME, ATOM14_AMBIGUOUS = _atom14_tables()


# --- PDB text -> atom37 arrays ---------------------------------------------------------------------------------------
def parse_pdb_string(pdb_str: str, chain_id: Optional[str] = None) -> Dict[str, np.ndarray]:
    """ATOM/HETATM columns -> {atom_positions [L,37,3] f64, atom_mask [L,37] f64, aatype [L] i64, residue_index [L] i64,
    chain_index [L] i64, b_factors [L,37] f64}, the dict `Protein.to_dict()` gives (protein.py:72-140).

    Same rules: one model only; unknown residue names -> 20 ('X'); atom names outside the 37 are dropped; a residue with
    no known atom is skipped; insertion codes are an error; coordinates go through fp32 (Biopython stores fp32
    coordinates); chain ids are numbered in sorted order (np.unique); for alternate locations the highest occupancy
    wins (Biopython's default selection), first seen on ties."""
    residues: Dict[tuple, dict] = {}
    order: List[tuple] = []
    n_models = 0
    ended = False
    for line in pdb_str.splitlines():
        rec = line[:6]
        if rec.startswith("MODEL"):
            n_models += 1
            if n_models > 1:
                raise ValueError("Only single model PDBs are supported. Found more than one MODEL record.")
            continue
        if rec.startswith("ENDMDL"):
            ended = True
            continue
        if rec not in ("ATOM  ", "HETATM"):
            continue
        if ended:
            raise ValueError("Only single model PDBs are supported. Found atoms after ENDMDL.")
        chain = line[21]
        if chain_id is not None and chain != chain_id:
            continue
        if line[26] != " ":
            raise ValueError(f"PDB contains an insertion code at chain {chain} and residue index {int(line[22:26])}. "
                             "These are not supported.")
        hetflag = "H" if rec == "HETATM" else " "
        key = (chain, hetflag, int(line[22:26]))
        res = residues.get(key)
        if res is None:
            res = residues[key] = {"resname": line[17:20].strip(), "atoms": {}}
            order.append(key)
        name = line[12:16].strip()
        if name not in ATOM_ORDER:
            continue
        try:
            occ = float(line[54:60])
        except ValueError:
            occ = 1.0
        try:
            bf = float(line[60:66])
        except ValueError:
            bf = 0.0
        xyz = np.array([line[30:38], line[38:46], line[46:54]], dtype=np.float32)
        prev = res["atoms"].get(name)
        if prev is None or occ > prev[1]:
            res["atoms"][name] = (xyz, occ, bf)
    # Biopython walks `for chain in model: for res in chain`: residues are grouped by chain in first-seen chain order (a
    # chain's HETATM / water records that appear after another chain's atoms still follow that chain's own residues)
    chain_rank: Dict[str, int] = {}
    for key in order:
        chain_rank.setdefault(key[0], len(chain_rank))
    order.sort(key=lambda k: chain_rank[k[0]])  # stable: first-appearance order inside a chain is kept
    pos, mask, aatype, resid, chains, bfs = [], [], [], [], [], []
    for key in order:
        res = residues[key]
        if not res["atoms"]:
            continue
        p = np.zeros((37, 3))
        m = np.zeros((37,))
        b = np.zeros((37,))
        for name, (xyz, _, bf) in res["atoms"].items():
            i = ATOM_ORDER[name]
            p[i], m[i], b[i] = xyz, 1.0, bf
        pos.append(p)
        mask.append(m)
        bfs.append(b)
        aatype.append(RESNAME_TO_IDX.get(res["resname"], 20))
        resid.append(key[2])
        chains.append(key[0])
    uniq = {c: n for n, c in enumerate(np.unique(chains))} if chains else {}
    return {"atom_positions": np.array(pos), "atom_mask": np.array(mask), "aatype": np.array(aatype, dtype=np.int64),
            "residue_index": np.array(resid, dtype=np.int64), "chain_index": np.array([uniq[c] for c in chains], dtype=np.int64),
            "b_factors": np.array(bfs)}


def parse_pdb(path: str, chain_id: Optional[str] = None) -> Dict[str, np.ndarray]:
    with open(path, "r") as f:
        return parse_pdb_string(f.read(), chain_id)


# --- atom37 -> frames / torsions -------------------------------------------------------------------------------------
def _frames_from_3_points(neg_x: torch.Tensor, origin: torch.Tensor, xy: torch.Tensor, eps: float = 1e-8):
    """Gram-Schmidt frame (rigid_utils.py:1236-1278): columns e0, e1, e2; returns (R [...,3,3], t [...,3]) in the input
    dtype — the caller rounds to fp32 where the reference's Rigid would."""
    e0 = origin - neg_x
    e1 = xy - origin
    e0 = e0 / torch.sqrt((e0[..., 0] * e0[..., 0] + e0[..., 1] * e0[..., 1] + e0[..., 2] * e0[..., 2]) + eps)[..., None]
    dot = e0[..., 0] * e1[..., 0] + e0[..., 1] * e1[..., 1] + e0[..., 2] * e1[..., 2]
    e1 = e1 - e0 * dot[..., None]
    e1 = e1 / torch.sqrt((e1[..., 0] * e1[..., 0] + e1[..., 1] * e1[..., 1] + e1[..., 2] * e1[..., 2]) + eps)[..., None]
    e2 = torch.stack([e0[..., 1] * e1[..., 2] - e0[..., 2] * e1[..., 1],
                      e0[..., 2] * e1[..., 0] - e0[..., 0] * e1[..., 2],
                      e0[..., 0] * e1[..., 1] - e0[..., 1] * e1[..., 0]], -1)
    return torch.stack([e0, e1, e2], -1), origin


def _to_4x4(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    out = t.new_zeros(*t.shape[:-1], 4, 4)
    out[..., :3, :3] = R
    out[..., :3, 3] = t
    out[..., 3, 3] = 1
    return out


def atom37_to_frames(aatype: torch.Tensor, pos: torch.Tensor, mask: torch.Tensor, eps: float = 1e-8) -> Dict[str, torch.Tensor]:
    """[L] i64, [L,37,3] f64, [L,37] f64 -> the five `rigidgroups_*` features (data_transforms.py:758-894).
    The reference's two composes with sign-flip rotations are column sign flips: group 0 negates the x and z axes
    (:840-846), an ambiguous group's alternative frame negates y and z (:859-880)."""
    aatype = aatype.clamp(max=20)
    dev = pos.device  # every table lookup / index below lives on the device of the inputs: the same code featurises on the GPU
    base = torch.as_tensor(GROUP_BASE_ATOMS, device=dev)[aatype]                      # [L,8,3] atom37 indices
    L = aatype.shape[0]
    ar = torch.arange(L, device=dev)[:, None, None]
    p = pos[ar, base]                                                                  # [L,8,3,3]
    R, t = _frames_from_3_points(p[..., 0, :], p[..., 1, :], p[..., 2, :], eps)
    R, t = R.float(), t.float()                                                        # Rigid forces fp32
    flip0 = torch.ones(8, 3, device=dev)
    flip0[0, 0] = flip0[0, 2] = -1.0
    R = R * flip0[None, :, None, :]
    group_exists = torch.as_tensor(GROUP_MASK, dtype=mask.dtype, device=dev)[aatype]
    gt_exists = mask[ar, base].min(-1)[0] * group_exists
    amb = torch.as_tensor(GROUP_AMBIGUOUS, dtype=mask.dtype, device=dev)[aatype]      # [L,8]
    alt_sign = torch.ones(L, 8, 3, device=dev)
    alt_sign[..., 1:] = (1.0 - 2.0 * amb.float())[..., None]
    return {"rigidgroups_gt_frames": _to_4x4(R, t), "rigidgroups_gt_exists": gt_exists, "rigidgroups_group_exists": group_exists,
            "rigidgroups_group_is_ambiguous": amb, "rigidgroups_alt_gt_frames": _to_4x4(R * alt_sign[..., None, :], t)}


def atom37_to_torsion_angles(aatype: torch.Tensor, pos: torch.Tensor, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """(sin, cos) of pre-omega, phi, psi, chi1..4 per residue (data_transforms.py:925-1090): the fourth atom of each
    dihedral in the frame of the first three; psi's sign flipped (:1060-1062); chain start has no omega / phi."""
    aatype = aatype.clamp(max=20)
    L = aatype.shape[0]
    prev_pos = torch.cat([pos.new_zeros(1, 37, 3), pos[:-1]], 0)
    prev_mask = torch.cat([mask.new_zeros(1, 37), mask[:-1]], 0)
    dev = pos.device
    quad = torch.empty(L, 7, 4, 3, dtype=pos.dtype, device=dev)
    quad[:, 0] = torch.cat([prev_pos[:, 1:3], pos[:, :2]], 1)                          # CA-, C-, N, CA
    quad[:, 1] = torch.cat([prev_pos[:, 2:3], pos[:, :3]], 1)                          # C-, N, CA, C
    quad[:, 2] = torch.cat([pos[:, :3], pos[:, 4:5]], 1)                               # N, CA, C, O
    chi_idx = torch.as_tensor(CHI_ATOM_IDX, device=dev)[aatype]                        # [L,4,4]
    ar = torch.arange(L, device=dev)[:, None, None]
    quad[:, 3:] = pos[ar, chi_idx]
    tmask = torch.empty(L, 7, dtype=mask.dtype, device=dev)
    tmask[:, 0] = prev_mask[:, 1] * prev_mask[:, 2] * mask[:, 0] * mask[:, 1]
    tmask[:, 1] = prev_mask[:, 2] * (mask[:, 0] * mask[:, 1] * mask[:, 2])
    tmask[:, 2] = (mask[:, 0] * mask[:, 1] * mask[:, 2]) * mask[:, 4]
    tmask[:, 3:] = torch.as_tensor(CHI_MASK, dtype=mask.dtype, device=dev)[aatype] * mask[ar, chi_idx].prod(-1)
    R, t = _frames_from_3_points(quad[..., 1, :], quad[..., 2, :], quad[..., 0, :], 1e-8)
    R, t = R.float(), t.float()
    # Rigid.invert() in fp32 (R^T, -(R^T t)), applied to the fp64 fourth atom (promotion), rigid_utils.py:1135-1145
    Rt = R.transpose(-1, -2)
    tinv = -1 * (Rt[..., :, 0] * t[..., None, 0] + Rt[..., :, 1] * t[..., None, 1] + Rt[..., :, 2] * t[..., None, 2])
    p4 = quad[..., 3, :]
    rel = Rt[..., :, 0] * p4[..., None, 0] + Rt[..., :, 1] * p4[..., None, 1] + Rt[..., :, 2] * p4[..., None, 2] + tinv
    sc = torch.stack([rel[..., 2], rel[..., 1]], -1)
    sc = sc / torch.sqrt((sc * sc).sum(-1, keepdim=True) + 1e-8)
    sc = sc * sc.new_tensor([1.0, 1.0, -1.0, 1.0, 1.0, 1.0, 1.0])[None, :, None]
    mirror = torch.cat([mask.new_ones(L, 3), 1.0 - 2.0 * torch.as_tensor(CHI_PI_PERIODIC, dtype=sc.dtype, device=dev)[aatype]], -1)
    return {"torsion_angles_sin_cos": sc, "alt_torsion_angles_sin_cos": sc * mirror[..., None], "torsion_angles_mask": tmask}


def pseudo_beta_and_atom14(aatype: torch.Tensor, pos: torch.Tensor, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Pseudo-beta (CB, CA for glycine; data_transforms.py:370-387) and the dense 14-atom view of the atom37 arrays with its
    renamed alternative for the residues whose atom names are interchangeable (:575-757)."""
    L, dev = aatype.shape[0], pos.device
    is_gly = aatype == RESNAME_TO_IDX["GLY"]
    ca, cb = ATOM_ORDER["CA"], ATOM_ORDER["CB"]
    out = {"pseudo_beta": torch.where(is_gly[:, None], pos[:, ca], pos[:, cb]), "pseudo_beta_mask": torch.where(is_gly, mask[:, ca], mask[:, cb])}
    a14_to_37 = torch.as_tensor(ATOM14_TO_37, device=dev)[aatype]
    exists14 = torch.as_tensor(ATOM14_MASK, device=dev)[aatype]
    ar = torch.arange(L, device=dev)[:, None]
    gt_exists = exists14 * mask[ar, a14_to_37]
    gt_pos = gt_exists[..., None] * pos[ar, a14_to_37]
    rename = torch.as_tensor(ATOM14_RENAME, dtype=mask.dtype, device=dev)[aatype]
    out.update({
        "atom14_atom_exists": exists14, "residx_atom14_to_atom37": a14_to_37, "residx_atom37_to_atom14": torch.as_tensor(ATOM37_TO_14, device=dev)[aatype],
        "atom37_atom_exists": torch.as_tensor(ATOM37_MASK, device=dev)[aatype], "atom14_gt_exists": gt_exists, "atom14_gt_positions": gt_pos,
        "atom14_alt_gt_positions": torch.einsum("rac,rab->rbc", gt_pos, rename), "atom14_alt_gt_exists": torch.einsum("ra,rab->rb", gt_exists, rename),
        "atom14_atom_is_ambiguous": torch.as_tensor(ATOM14_AMBIGUOUS, dtype=mask.dtype, device=dev)[aatype],
    })
    return out


class ProteinFeatureTransform:
    """Same constructor keys as the reference's class (configs/data/sampling.yaml:8-13; dataset.py:26-47) and the same
    order of operations in `__call__` (:49-68)."""

    def __init__(self, unit: Optional[str] = "angstrom", truncate_length: Optional[int] = None, strip_missing_residues: bool = True,
                 recenter_and_scale: bool = True, eps: float = 1e-8, device=None):
        """`device` (not a reference key): where the geometric features (frames, torsions, pseudo-beta, atom14 views) are
        computed and returned, e.g. "cuda" to featurise straight into the sampler's device; None = CPU like the reference."""
        self.device = device
        if unit == "angstrom":
            self.coordinate_scale = 1.0
        elif unit in ("nm", "nanometer"):
            self.coordinate_scale = 0.1
        else:
            raise ValueError(f"Invalid unit: {unit}")
        if truncate_length is not None:
            assert truncate_length > 0, f"Invalid truncate_length: {truncate_length}"
        self.truncate_length = truncate_length
        self.strip_missing_residues = strip_missing_residues
        self.recenter_and_scale = recenter_and_scale
        self.eps = eps

    def __call__(self, chain_feats: Dict[str, np.ndarray]) -> Dict[str, torch.Tensor]:
        f = dict(chain_feats)
        seq_mask = f["atom_mask"][:, CA_IDX]
        f.update(seq_mask=seq_mask, residue_mask=seq_mask, residue_idx=f["residue_index"] - np.min(f["residue_index"]),
                 fixed_mask=np.zeros_like(seq_mask), sc_ca_t=np.zeros(seq_mask.shape + (3,)))
        if self.strip_missing_residues:                      # drop unknown residues at both ends (:84-91)
            known = np.where(f["aatype"] != 20)[0]
            f = {k: v[known.min():known.max() + 1] for k, v in f.items()}
        if self.truncate_length is not None and f["aatype"].shape[0] > self.truncate_length:
            start = np.random.randint(0, f["aatype"].shape[0] - self.truncate_length + 1)   # same RNG call as :98
            f = {k: v[start:start + self.truncate_length] for k, v in f.items()}
        if self.recenter_and_scale:                          # :114-122
            centre = np.sum(f["atom_positions"][:, CA_IDX], axis=0) / (np.sum(f["seq_mask"]) + self.eps)
            f["atom_positions"] = (f["atom_positions"] - centre[None, None, :]) * self.coordinate_scale * f["atom_mask"][..., None]
        t = {k: torch.as_tensor(v) for k, v in f.items()}
        if self.device is not None:
            t = {k: v.to(self.device) for k, v in t.items()}
        t["aatype"] = t["aatype"].long()
        t["atom_positions"] = t["atom_positions"].double()
        t["atom_mask"] = t["atom_mask"].double()
        t.update(atom37_to_frames(t["aatype"], t["atom_positions"], t["atom_mask"]))
        t.update(atom37_to_torsion_angles(t["aatype"], t["atom_positions"], t["atom_mask"]))
        t["backbone_rigid_tensor"] = t["rigidgroups_gt_frames"][..., 0, :, :]
        t["backbone_rigid_mask"] = t["rigidgroups_gt_exists"][..., 0]
        t["chi_angles_sin_cos"] = t["torsion_angles_sin_cos"][..., 3:, :].to(t["atom_mask"].dtype)
        t["chi_mask"] = t["torsion_angles_mask"][..., 3:].to(t["atom_mask"].dtype)
        t.update(pseudo_beta_and_atom14(t["aatype"].clamp(max=20), t["atom_positions"], t["atom_mask"]))
        return t


class SamplingPDBDataset:
    """Directory of .pdb files, sorted, optional accession filter (dataset.py:186-247, 304-320)."""

    def __init__(self, path_to_dataset: str, training: bool = False, suffix: str = ".pdb", transform: Optional[ProteinFeatureTransform] = None,
                 accession_code_fillter: Optional[Sequence[str]] = None):
        path_to_dataset = os.path.expanduser(path_to_dataset)
        assert os.path.isdir(path_to_dataset), f"Invalid path (expected to be directory): {path_to_dataset}"
        suffix = suffix if suffix.startswith(".") else "." + suffix
        assert suffix == ".pdb", f"Invalid suffix: {suffix}"
        data = sorted(glob.glob(os.path.join(path_to_dataset, "*" + suffix)))
        assert len(data) > 0, f"No {suffix} file found in '{path_to_dataset}'"
        if accession_code_fillter and len(accession_code_fillter) > 0:
            data = [p for p in data if os.path.splitext(os.path.basename(p))[0] in set(accession_code_fillter)]
        self.data, self.transform, self.training = data, transform, training

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, idx: int):
        path = self.data[idx]
        obj = parse_pdb(path)
        if self.transform is not None:
            obj = self.transform(obj)
        obj["accession_code"] = os.path.splitext(os.path.basename(path))[0]
        return obj


def collate(raw_batch: Sequence[Dict[str, object]], pad_v: float = 0.0) -> Dict[str, object]:
    """`BatchTensorConverter` (protein_datamodule.py:9-57): tensors are zero-padded to the longest sample and stacked,
    everything else is returned as a list."""
    keys = [k for k, v in raw_batch[0].items() if torch.is_tensor(v)]
    out: Dict[str, object] = {}
    for k in keys:
        xs = [d[k] for d in raw_batch]
        if len({x.dim() for x in xs}) != 1:
            raise RuntimeError(f"Samples has varying dimensions: {[x.dim() for x in xs]}")
        shape = [max(s) for s in zip(*[x.shape for x in xs])]
        res = torch.full((len(xs), *shape), pad_v, dtype=xs[0].dtype)
        for i, x in enumerate(xs):
            res[(i,) + tuple(slice(0, n) for n in x.shape)] = x
        out[k] = res
    for k in raw_batch[0]:
        if k not in keys:
            out[k] = [d[k] for d in raw_batch]
    return out


def featurize_pdb(path: str, transform: Optional[ProteinFeatureTransform] = None) -> Dict[str, object]:
    """One PDB file -> the batch-of-one dict `predict_step` takes (sampling config: no stripping, no recentring)."""
    transform = transform or ProteinFeatureTransform(truncate_length=None, strip_missing_residues=False, recenter_and_scale=False, eps=1e-8)
    obj = transform(parse_pdb(path))
    obj["accession_code"] = os.path.splitext(os.path.basename(path))[0]
    return collate([obj])
