"""Golden vectors for str2str_b200.featurize, produced by the UNMODIFIED reference in the build container:

  * PDB text written by the reference's own writer (src/common/protein.py:to_pdb) from seeded arrays — the parser must
    recover exactly those arrays (Biopython, which the reference parses with, is not installed; its writer is the pin);
  * the reference's `ProteinFeatureTransform` (src/data/components/dataset.py:26-143 -> data_transforms.atom37_to_frames,
    atom37_to_torsion_angles, get_backbone_frames) applied to those arrays, for the sampling config and for the
    strip + recentre config.

    python tests/golden/make_golden_featurize.py

Also asserts (not stored: the bundled structures are the reference's data) that every .pdb under /root/reference/data
featurises identically through both implementations.
"""
import glob
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import refshim  # noqa: E402

refshim.install()
for name, attrs in (("Bio", {}), ("Bio.PDB", {"PDBParser": object}), ("biotite", {}), ("biotite.structure", {}),
                    ("biotite.structure.io", {}), ("biotite.structure.io.pdb", {"PDBFile": object})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
from src.common import protein as ref_protein  # noqa: E402
from src.common import residue_constants as rc  # noqa: E402
from src.data.components.dataset import ProteinFeatureTransform as RefTransform  # noqa: E402

from str2str_b200 import featurize as F  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KEYS = ["aatype", "residue_index", "residue_idx", "chain_index", "atom_positions", "atom_mask", "seq_mask", "residue_mask", "fixed_mask",
        "sc_ca_t", "rigidgroups_gt_frames", "rigidgroups_gt_exists", "rigidgroups_group_exists", "rigidgroups_group_is_ambiguous",
        "rigidgroups_alt_gt_frames", "torsion_angles_sin_cos", "alt_torsion_angles_sin_cos", "torsion_angles_mask",
        "backbone_rigid_tensor", "backbone_rigid_mask", "b_factors", "chi_angles_sin_cos", "chi_mask", "pseudo_beta", "pseudo_beta_mask",
        "atom14_atom_exists", "residx_atom14_to_atom37", "residx_atom37_to_atom14", "atom37_atom_exists", "atom14_gt_exists",
        "atom14_gt_positions", "atom14_alt_gt_positions", "atom14_alt_gt_exists", "atom14_atom_is_ambiguous"]


def check_tables():
    assert list(rc.restypes) == list(F.RESTYPES) and [rc.restype_1to3[r] for r in rc.restypes] == list(F.RESNAMES)
    assert list(rc.atom_types) == list(F.ATOM_TYPES)
    from src.common.data_transforms import get_chi_atom_indices

    assert np.array_equal(np.array(get_chi_atom_indices()), F.CHI_ATOM_IDX)
    assert np.array_equal(np.array(list(rc.chi_angles_mask) + [[0.0] * 4]), F.CHI_MASK)
    assert np.array_equal(np.array(rc.chi_pi_periodic), F.CHI_PI_PERIODIC)


def synthetic_protein(L, seed, unk_ends=0, missing=True, two_chains=False):
    """Random-walk CA trace with every heavy atom of each residue type placed near its CA (geometry need not be
    chemical for a featurisation pin); coordinates on the PDB 0.001 grid."""
    rng = np.random.default_rng(seed)
    aatype = rng.integers(0, 20, L)
    aatype[:: max(2, L // 7)] = [3, 6, 13, 18, 1, 11, 7, 0][: len(aatype[:: max(2, L // 7)])]   # ASP GLU PHE TYR ARG LYS GLY ALA
    if unk_ends:
        aatype[:unk_ends] = 20
        aatype[-unk_ends:] = 20
    step = rng.normal(size=(L, 3))
    ca = np.cumsum(3.8 * step / np.linalg.norm(step, axis=-1, keepdims=True), 0)
    pos = np.zeros((L, 37, 3))
    mask = np.zeros((L, 37))
    for i, a in enumerate(aatype):
        names = ["N", "CA", "C", "O"] if a == 20 else [n for n in rc.restype_name_to_atom14_names[rc.restype_1to3[rc.restypes[a]]] if n]
        for n in names:
            j = rc.atom_order[n]
            pos[i, j] = ca[i] + (0 if n == "CA" else rng.normal(0, 1.5, 3))
            mask[i, j] = 1.0
    if missing:                                                # a missing side-chain tip, a missing O, a missing CA
        for i in rng.choice(np.arange(1, L - 1), size=min(3, L - 2), replace=False):
            tip = np.where(mask[i] > 0)[0][-1]
            mask[i, tip] = 0
        mask[L // 2, rc.atom_order["O"]] = 0
        mask[L // 3, rc.atom_order["CA"]] = 0
    # Biopython (the reference's parser) holds coordinates as fp32: the arrays a parsed file yields are fp32-rounded
    pos = np.round(pos, 3).astype(np.float32).astype(np.float64) * mask[..., None]
    residue_index = np.arange(L) + 5
    residue_index[L // 2:] += 7                                 # chain break
    chain_index = np.zeros(L, np.int64)
    if two_chains:
        chain_index[2 * L // 3:] = 1
    b = np.round(rng.uniform(10, 90, (L, 37)), 2) * mask
    return dict(atom_positions=pos, atom_mask=mask, aatype=aatype, residue_index=residue_index, chain_index=chain_index, b_factors=b)


def run_ref(raw, **kw):
    out = RefTransform(**kw)({k: np.array(v) for k, v in raw.items()})
    assert sorted(out.keys()) == sorted(KEYS), sorted(set(out.keys()) ^ set(KEYS))   # every tensor of the reference transform is pinned
    return {k: out[k].numpy() for k in KEYS}


def compare(a, b, where):
    for k in KEYS:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape and x.dtype == y.dtype, (where, k, x.shape, y.shape, x.dtype, y.dtype)
        err = float(np.abs(x.astype(np.float64) - y.astype(np.float64)).max()) if x.size else 0.0
        assert err <= 2e-6, (where, k, err)


def main():
    check_tables()
    cases = {"small_sampling": (dict(L=24, seed=1), dict(truncate_length=None, strip_missing_residues=False, recenter_and_scale=False, eps=1e-8)),
             "unk_ends_strip_recentre": (dict(L=30, seed=2, unk_ends=2), dict(truncate_length=None, strip_missing_residues=True, recenter_and_scale=True, eps=1e-8)),
             "two_chains": (dict(L=40, seed=3, two_chains=True), dict(truncate_length=None, strip_missing_residues=False, recenter_and_scale=False, eps=1e-8))}
    for name, (skw, tkw) in cases.items():
        raw = synthetic_protein(**skw)
        text = ref_protein.to_pdb(ref_protein.Protein(**raw))
        ours = {k: v.numpy() for k, v in F.ProteinFeatureTransform(**tkw)(F.parse_pdb_string(text)).items() if k in KEYS}
        gold = run_ref(raw, **tkw)
        compare(ours, gold, name)
        np.savez_compressed(os.path.join(OUT, f"featurize_{name}.npz"), pdb_text=np.array(text), transform_kwargs=np.array(repr(tkw)),
                            **{f"raw_{k}": v for k, v in raw.items()}, **{f"out_{k}": v for k, v in gold.items()})
        print(f"wrote featurize_{name}.npz (L={raw['aatype'].shape[0]})")
    for path in sorted(glob.glob("/root/reference/data/*/*.pdb")):
        raw = F.parse_pdb(path)
        tkw = dict(truncate_length=None, strip_missing_residues=False, recenter_and_scale=False, eps=1e-8)
        ours = {k: v.numpy() for k, v in F.ProteinFeatureTransform(**tkw)(raw).items() if k in KEYS}
        compare(ours, run_ref(raw, **tkw), path)
        # the reference writer reproduces the file's ATOM lines from the parsed arrays (parser pin on real structures)
        text = ref_protein.to_pdb(ref_protein.Protein(**raw))
        atoms_in = [ln[:66].rstrip() for ln in open(path) if ln.startswith("ATOM")]
        atoms_out = [ln[:66].rstrip() for ln in text.splitlines() if ln.startswith("ATOM")]
        assert len(atoms_in) == len(atoms_out), (path, len(atoms_in), len(atoms_out))
        same = sum(a[12:] == b[12:] for a, b in zip(atoms_in, atoms_out))
        print(f"{os.path.basename(path)}: L={raw['aatype'].shape[0]} features match the reference; "
              f"{same}/{len(atoms_in)} ATOM lines reproduced by the reference writer from the parsed arrays")


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    main()
