"""str2str_b200.featurize (the step before the path) against goldens made by the unmodified reference
(tests/golden/make_golden_featurize.py): PDB text from the reference's writer -> parser recovers the arrays exactly;
ProteinFeatureTransform -> frames / torsions / masks within 2e-6 (float), integers and masks exact."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from str2str_b200 import featurize as F

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "featurize_*.npz")))
TOL = 2e-6  # fp32 frames / fp64 torsions built from fp32-rounded frames: reordering noise only


def test_goldens_present():
    assert len(CASES) == 3


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[10:-4] for p in CASES])
def test_parser_recovers_reference_written_arrays(path):
    g = np.load(path)
    raw = F.parse_pdb_string(str(g["pdb_text"]))
    for k in ("atom_positions", "atom_mask", "aatype", "residue_index", "chain_index", "b_factors"):
        ref = g[f"raw_{k}"]
        assert raw[k].shape == ref.shape, k
        assert np.array_equal(raw[k], ref), k            # coordinates sit on the fp32-rounded 0.001 grid: exact


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[10:-4] for p in CASES])
def test_transform_matches_reference(path):
    g = np.load(path)
    kw = ast.literal_eval(str(g["transform_kwargs"]))
    out = F.ProteinFeatureTransform(**kw)(F.parse_pdb_string(str(g["pdb_text"])))
    keys = [k[4:] for k in g.files if k.startswith("out_")]
    assert len(keys) == 34   # every tensor of the reference transform
    for k in keys:
        ref, got = g[f"out_{k}"], out[k].numpy()
        assert got.shape == ref.shape and got.dtype == ref.dtype, (k, got.shape, ref.shape, got.dtype, ref.dtype)
        if np.issubdtype(ref.dtype, np.integer) or k.endswith("mask") or k.endswith("exists") or k.endswith("ambiguous"):
            assert np.array_equal(got, ref), k
        else:
            assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= TOL, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[10:-4] for p in CASES])
def test_transform_on_device_matches_reference(path):
    """The same transform with device="cuda": every geometric feature is computed by batched torch ops on the GPU and lands
    where the sampler needs it; same goldens of the unmodified reference, same tolerances."""
    g = np.load(path)
    kw = ast.literal_eval(str(g["transform_kwargs"]))
    out = F.ProteinFeatureTransform(device="cuda", **kw)(F.parse_pdb_string(str(g["pdb_text"])))
    keys = [k[4:] for k in g.files if k.startswith("out_")]
    assert len(keys) == 34
    for k in keys:
        assert out[k].is_cuda, k
        ref, got = g[f"out_{k}"], out[k].cpu().numpy()
        assert got.shape == ref.shape and got.dtype == ref.dtype, (k, got.shape, ref.shape, got.dtype, ref.dtype)
        if np.issubdtype(ref.dtype, np.integer) or k.endswith("mask") or k.endswith("exists") or k.endswith("ambiguous"):
            assert np.array_equal(got, ref), k
        else:
            assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= TOL, k


def _line(serial, name, resname, chain, resseq, xyz, icode=" ", altloc=" ", occ=1.0, rec="ATOM"):
    nm = name if len(name) == 4 else f" {name}"
    return (f"{rec:<6}{serial:>5} {nm:<4}{altloc}{resname:>3} {chain}{resseq:>4}{icode}   "
            f"{xyz[0]:>8.3f}{xyz[1]:>8.3f}{xyz[2]:>8.3f}{occ:>6.2f}{20.0:>6.2f}          {name[0]:>2}")


def test_parser_edge_cases():
    lines = [_line(1, "N", "ALA", "B", 3, (1, 2, 3)), _line(2, "CA", "ALA", "B", 3, (2, 2, 3), altloc="A", occ=0.4),
             _line(3, "CA", "ALA", "B", 3, (9, 9, 9), altloc="B", occ=0.6), _line(4, "XX", "ALA", "B", 3, (0, 0, 1)),
             _line(5, "N", "MSE", "A", 1, (4, 5, 6)), _line(6, "H1", "GLY", "A", 2, (4, 5, 6))]
    raw = F.parse_pdb_string("\n".join(lines))
    assert raw["aatype"].tolist() == [0, 20]                    # unknown residue name -> 20; the residue with no known atom is skipped
    assert raw["chain_index"].tolist() == [1, 0]                # file order kept, ids numbered in sorted order
    assert raw["residue_index"].tolist() == [3, 1]
    assert raw["atom_positions"][0, F.CA_IDX].tolist() == [9.0, 9.0, 9.0]   # highest occupancy wins
    assert raw["atom_mask"][0].sum() == 2
    only_a = F.parse_pdb_string("\n".join(lines), chain_id="A")
    assert only_a["aatype"].tolist() == [20]
    with pytest.raises(ValueError, match="insertion code"):
        F.parse_pdb_string(_line(1, "N", "ALA", "A", 3, (1, 2, 3), icode="A"))
    two = "MODEL     1\n" + lines[0] + "\nENDMDL\nMODEL     2\n" + lines[0] + "\nENDMDL\n"
    with pytest.raises(ValueError, match="single model"):
        F.parse_pdb_string(two)
    with pytest.raises(ValueError, match="Invalid unit"):
        F.ProteinFeatureTransform(unit="parsec")


def test_collate_pads_and_featurize_pdb_feeds_the_sampler(tmp_path):
    g = np.load(CASES[0])
    p = tmp_path / "prot_a.pdb"
    p.write_text(str(g["pdb_text"]))
    batch = F.featurize_pdb(str(p))
    L = g["raw_aatype"].shape[0]
    assert batch["accession_code"] == ["prot_a"]
    want = {"aatype": (torch.int64, (1, L)), "residue_idx": (torch.int64, (1, L)), "residue_mask": (torch.float64, (1, L)),
            "fixed_mask": (torch.float64, (1, L)), "sc_ca_t": (torch.float64, (1, L, 3)), "rigidgroups_gt_frames": (torch.float32, (1, L, 8, 4, 4)),
            "torsion_angles_sin_cos": (torch.float64, (1, L, 7, 2)), "chain_index": (torch.int64, (1, L)), "residue_index": (torch.int64, (1, L))}
    for k, (dt, shp) in want.items():                           # the dtypes of SURVEY §8b "Batch dict (producer contract)"
        assert batch[k].dtype == dt and tuple(batch[k].shape) == shp, k
    ds = F.SamplingPDBDataset(str(tmp_path), transform=F.ProteinFeatureTransform(strip_missing_residues=False, recenter_and_scale=False))
    assert len(ds) == 1 and ds[0]["accession_code"] == "prot_a"
    short = {k: (v[:5] if torch.is_tensor(v) else v) for k, v in ds[0].items()}
    both = F.collate([ds[0], short])
    assert both["aatype"].shape == (2, L) and both["aatype"][1, 5:].abs().sum() == 0
    assert both["rigidgroups_gt_frames"].shape == (2, L, 8, 4, 4) and both["accession_code"] == ["prot_a", "prot_a"]
