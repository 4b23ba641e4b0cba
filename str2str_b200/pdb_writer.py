"""Vectorised PDB writer for sampled ensembles — the step right after the hot path.

Byte-identical to the reference's `atom37_to_pdb` (src/common/pdb_utils.py:205-252) -> `protein.to_pdb`
(src/common/protein.py:152-234), which formats every atom with a Python f-string (≈650 k lines for 512 decoys of a
256-residue chain).  Here the per-line text that does not depend on the coordinates (record, serial, atom / residue names,
chain, residue number, occupancy, B-factor, element, TER / MODEL / ENDMDL lines) is built once per atom mask, and the
three `%8.3f` coordinate fields of all models are written into a `[models, lines, 81]` byte buffer with integer
arithmetic: for float32 input `x * 1000` is exact in float64, so `rint` reproduces Python's correctly-rounded
(half-to-even) formatting, including `-0.000` for negative values that round to zero.  Values that do not fit their
column width (|x| >= 10000 or x <= -1000), non-finite values, float64 positions and non-zero B-factors take the
reference's own formatting expression line by line, so the output is identical there too (just not fast).
Pinned by tests/golden/pdb_*.npz, generated from the unmodified reference writer (tests/golden/make_golden_pdb.py).
"""
from __future__ import annotations

import os
import re
from typing import Optional

import numpy as np

# src/common/residue_constants.py: atom_types (37), restypes (20) + 'X', restype_1to3
ATOM_TYPES = ["N", "CA", "C", "CB", "O", "CG", "CG1", "CG2", "OG", "OG1", "SG", "CD", "CD1", "CD2", "ND1", "ND2", "OD1", "OD2",
              "SD", "CE", "CE1", "CE2", "CE3", "NE", "NE1", "NE2", "OE1", "OE2", "CH2", "NH1", "NH2", "OH", "CZ", "CZ2", "CZ3",
              "NZ", "OXT"]
RESTYPES_3 = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER",
              "THR", "TRP", "TYR", "VAL", "UNK"]
PDB_CHAIN_IDS = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"
_CB = ATOM_TYPES.index("CB")
_GLY = RESTYPES_3.index("GLY")
_W = 81  # 80 columns + newline


def _squeeze(x):
    x = np.asarray(x)
    return np.squeeze(x) if x.shape[0] == 1 and x.ndim > 1 else x


def _atom_line(serial, atom_name, res3, chain, resi, pos, b_factor):
    """The reference's formatting expression (protein.py:199-214), used for the static parts and for fallbacks."""
    name = atom_name if len(atom_name) == 4 else f" {atom_name}"
    return (f"{'ATOM':<6}{serial:>5} {name:<4}{'':>1}{res3:>3} {chain:>1}{resi:>4}{'':>1}   "
            f"{pos[0]:>8.3f}{pos[1]:>8.3f}{pos[2]:>8.3f}{1.00:>6.2f}{b_factor:>6.2f}          {atom_name[0]:>2}{'':>2}")


def _ter_line(serial, res3, chain, resi):
    return f"{'TER':<6}{serial:>5}      {res3:>3} {chain:>1}{resi:>4}"


def _fixed3(x: np.ndarray):
    """'%8.3f' of float32 values as uint8 [..., 8]; second result marks the values that need the slow path."""
    xd = x.astype(np.float64)
    bad = ~np.isfinite(xd)
    q = np.rint(np.abs(np.where(bad, 0.0, xd)) * 1000.0).astype(np.int64)
    neg = np.signbit(xd)
    ip, fr = q // 1000, q % 1000
    bad |= (ip >= 10000) | (neg & (ip >= 1000))
    out = np.full(x.shape + (8,), ord(" "), np.uint8)
    out[..., 7] = 48 + fr % 10
    out[..., 6] = 48 + (fr // 10) % 10
    out[..., 5] = 48 + fr // 100
    out[..., 4] = ord(".")
    out[..., 3] = 48 + ip % 10
    nd = np.ones(x.shape, np.int64)  # integer digits written so far
    for k, p in ((2, 10), (1, 100), (0, 1000)):
        has = ip >= p
        out[..., k] = np.where(has, 48 + (ip // p) % 10, out[..., k])
        nd += has
    sign_col = 3 - nd  # column left of the most significant digit
    idx = np.nonzero(neg & ~bad)
    out[idx + (sign_col[idx],)] = ord("-")
    return out, bad


class _Template:
    """Static text of one model for one atom mask: [n_lines, 81] bytes + where the coordinates go."""

    def __init__(self, mask, aatype, residue_index, chain_index, b_factors):
        L = mask.shape[0]
        res3 = [RESTYPES_3[a] if 0 <= a < 21 else "UNK" for a in aatype]
        sel = mask.copy()
        sel[aatype == _GLY, _CB] = False  # "skip CB for GLY" (protein.py:192-194)
        lines, self.atom_line, self.atom_res, self.atom_type = [b"MODEL"], [], [], []
        serial = 1
        last_chain = chain_index[0]
        for i in range(L):
            if last_chain != chain_index[i]:
                lines.append(_ter_line(serial, res3[i - 1], PDB_CHAIN_IDS[chain_index[i - 1]], residue_index[i - 1]).encode())
                last_chain = chain_index[i]
                serial += 1
            chain = PDB_CHAIN_IDS[chain_index[i]]
            for k in np.nonzero(sel[i])[0]:
                self.atom_line.append(len(lines))
                self.atom_res.append(i)
                self.atom_type.append(int(k))
                lines.append(_atom_line(serial, ATOM_TYPES[k], res3[i], chain, residue_index[i], (0.0, 0.0, 0.0), b_factors[i, k]).encode())
                serial += 1
        lines.append(_ter_line(serial, res3[-1], PDB_CHAIN_IDS[chain_index[-1]], residue_index[-1]).encode())
        lines.append(b"ENDMDL")
        self.text = np.full((len(lines), _W), ord(" "), np.uint8)
        self.text[:, 80] = ord("\n")
        for n, ln in enumerate(lines):
            self.text[n, : len(ln)] = np.frombuffer(ln, np.uint8)
        self.atom_line = np.asarray(self.atom_line, np.int64)
        self.atom_res = np.asarray(self.atom_res, np.int64)
        self.atom_type = np.asarray(self.atom_type, np.int64)
        self.meta = (res3, residue_index, chain_index, b_factors)


def to_pdb_bytes(atom_positions: np.ndarray, aatype=None, b_factors=None, chain_index=None, residue_index=None) -> bytes:
    """All models of `atom_positions` ([M, L, 37, 3] or [L, 37, 3]) as the reference writes them, ending in 'END'."""
    pos = np.asarray(atom_positions)
    if pos.ndim == 3:
        pos = pos[None]
    if pos.ndim != 4 or pos.shape[-2:] != (37, 3):
        raise ValueError(f"Invalid positions shape {np.asarray(atom_positions).shape}")
    M, L = pos.shape[:2]
    residue_index = (np.arange(L) + 1 if residue_index is None else _squeeze(residue_index)).astype(int)
    chain_index = (np.zeros(L) if chain_index is None else _squeeze(chain_index)).astype(int)
    b_factors = np.zeros([L, 37]) if b_factors is None else _squeeze(b_factors)
    aatype = (np.zeros(L, dtype=int) if aatype is None else _squeeze(aatype)).astype(int)
    if np.any(aatype > 20):
        raise ValueError("Invalid aatypes.")
    if np.any(chain_index >= len(PDB_CHAIN_IDS)):
        raise ValueError(f"The PDB format supports at most {len(PDB_CHAIN_IDS)} chains.")
    masks = np.sum(np.abs(pos), axis=-1) > 1e-7  # [M, L, 37] (pdb_utils.py:232)
    fast = pos.dtype == np.float32
    chunks, templates = [], {}
    # models that share an atom mask share the static text; in practice all of them do
    keys = [m.tobytes() for m in masks] if not (masks == masks[0]).all() else [b""] * M
    for key in dict.fromkeys(keys):
        members = np.asarray([i for i, k in enumerate(keys) if k == key])
        tpl = templates[key] = _Template(masks[members[0]], aatype, residue_index, chain_index, b_factors)
        buf = np.broadcast_to(tpl.text, (len(members),) + tpl.text.shape).copy()
        xyz = pos[members][:, tpl.atom_res, tpl.atom_type]  # [m, n_atoms, 3]
        if fast:
            chars, bad = _fixed3(xyz)
            buf[:, tpl.atom_line, 30:54] = chars.reshape(len(members), -1, 24)
        else:
            bad = np.ones(xyz.shape, bool)
        chunks.append((members, buf, tpl, xyz, bad.any(-1)))
    out = [None] * M
    for members, buf, tpl, xyz, bad in chunks:
        res3, resi, chain, bf = tpl.meta
        for n, m in enumerate(members):
            head = f"MODEL     {m + 1}".encode()
            buf[n, 0, : len(head)] = np.frombuffer(head, np.uint8)
            if not bad[n].any():
                out[m] = buf[n].tobytes()
                continue
            lines = [bytes(row[:80]).decode() for row in buf[n]]  # slow path: re-format the offending lines like the reference
            for a in np.nonzero(bad[n])[0]:
                i, k, ln = tpl.atom_res[a], tpl.atom_type[a], tpl.atom_line[a]
                serial = int(lines[ln][6:11])
                lines[ln] = _atom_line(serial, ATOM_TYPES[k], res3[i], PDB_CHAIN_IDS[chain[i]], resi[i], xyz[n, a], bf[i, k]).ljust(80)
            out[m] = ("\n".join(lines) + "\n").encode()
    return b"".join(out) + b"END"


def atom37_to_pdb(save_to: str, atom_positions: np.ndarray, aatype: Optional[np.ndarray] = None, b_factors: Optional[np.ndarray] = None,
                  chain_index: Optional[np.ndarray] = None, residue_index: Optional[np.ndarray] = None, overwrite: bool = False,
                  no_indexing: bool = True):
    """Drop-in for the reference's `atom37_to_pdb` (same arguments, same file bytes, same return value)."""
    if overwrite:
        max_existing_idx = 0
    else:
        file_dir = os.path.dirname(save_to)
        file_name = os.path.basename(save_to).strip(".pdb")
        existing_files = [x for x in os.listdir(file_dir) if file_name in x]
        max_existing_idx = max([int(re.findall(r"_(\d+).pdb", x)[0]) for x in existing_files if re.findall(r"_(\d+).pdb", x)] + [0])
    if not no_indexing:
        save_to = save_to.replace(".pdb", "") + f"_{max_existing_idx + 1}.pdb"
    pos = np.asarray(atom_positions)
    if pos.ndim not in (3, 4):
        raise ValueError(f"Invalid positions shape {pos.shape}")
    data = to_pdb_bytes(pos, aatype=aatype, b_factors=b_factors, chain_index=chain_index, residue_index=residue_index)
    with open(save_to, "wb") as f:
        f.write(data)
    return save_to
