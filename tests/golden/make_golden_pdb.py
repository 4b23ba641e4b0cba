"""Golden PDB text for str2str_b200.pdb_writer, produced by the UNMODIFIED reference writer
(src/common/pdb_utils.py:atom37_to_pdb -> src/common/protein.py:to_pdb) in the build container.

    python tests/golden/make_golden_pdb.py

The reference writer needs Bio.PDB / biotite / tqdm only at import time (parsers, progress bars); empty stub modules are
enough because the writing path is pure string formatting.  Inputs are seeded and stored next to the text they produce.
"""
import os
import sys
import tempfile
import types

import numpy as np

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
if "tqdm" not in sys.modules:
    try:
        import tqdm  # noqa: F401
    except ImportError:
        m = types.ModuleType("tqdm")
        m.tqdm = lambda x, **k: x
        sys.modules["tqdm"] = m
from src.common.pdb_utils import atom37_to_pdb  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, n_models, L, seed, with_aatype, chain_break=None, residue_offset=None, gly=False, big=False):
    rng = np.random.default_rng(seed)
    pos = np.zeros((n_models, L, 37, 3), np.float32)
    bb = rng.normal(0, 12.0 if not big else 400.0, (n_models, L, 5, 3)).astype(np.float32)
    pos[:, :, [0, 1, 2, 4, 3]] = bb                      # N, CA, C, O, CB as compute_backbone fills them
    pos[0, 0, 1] = [-0.0004, 0.0005, -1.0005]            # negative zero after rounding, ties
    pos[-1, -1, 4] = [999.9994, -99.9996, 0.0015]
    if L > 3:
        pos[:, 3, 3] = 0.0                               # a masked atom (all-zero coordinates are skipped)
    kw = {}
    aatype = None
    if with_aatype:
        aatype = rng.integers(0, 20, (1, L))
        if gly:
            aatype[0, ::3] = 7                           # glycine: its CB line is dropped
        kw["aatype"] = aatype
    if chain_break is not None:
        ci = np.zeros((1, L), np.int64)
        ci[0, chain_break:] = 1
        kw["chain_index"] = ci
    if residue_offset is not None:
        ri = np.arange(L)[None] + residue_offset
        if chain_break is not None:
            ri[0, chain_break:] += 17
        kw["residue_index"] = ri
    pos_in = pos[0] if n_models == 0 else pos
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "x.pdb")
        atom37_to_pdb(save_to=path, atom_positions=pos_in, overwrite=True, **kw)
        text = open(path).read()
    np.savez_compressed(os.path.join(OUT, f"pdb_{name}.npz"), atom_positions=pos_in, text=np.frombuffer(text.encode(), np.uint8),
                        **{k: v for k, v in kw.items()})
    print(name, len(text), "bytes", text.count("\n"), "lines")


def merged_case():
    """merge_pdbfiles (pdb_utils.py:31-82) over two multi-model goldens and one single-model file without MODEL records."""
    from src.common.pdb_utils import merge_pdbfiles

    names = ["pdb_default_L12_m3.npz", "pdb_aatype_gly_L20_m2.npz"]
    with tempfile.TemporaryDirectory() as d:
        files = []
        for n in names:
            p = os.path.join(d, n[:-4] + ".pdb")
            open(p, "wb").write(bytes(np.load(os.path.join(OUT, n))["text"]))
            files.append(p)
        txt = bytes(np.load(os.path.join(OUT, names[0]))["text"]).decode().split("\n")
        first = []
        for ln in (x for x in txt[1:] if x.startswith("ATOM") or x.startswith("TER")):
            first.append(ln)
            if ln.startswith("TER"):
                break
        single = os.path.join(d, "single.pdb")
        open(single, "w").write("\n".join(first) + "\nEND\n")
        files.append(single)
        out = os.path.join(d, "o", "merged.pdb")
        merge_pdbfiles(files, out, verbose=False)
        np.savez_compressed(os.path.join(OUT, "pdb_merged.npz"), inputs=np.array(names),
                            single=np.frombuffer(open(single, "rb").read(), np.uint8), text=np.frombuffer(open(out, "rb").read(), np.uint8))


if __name__ == "__main__":
    case("default_L12_m3", 3, 12, 1, False)
    case("aatype_gly_L20_m2", 2, 20, 2, True, gly=True)
    case("chains_resid_L16_m2", 2, 16, 3, True, chain_break=9, residue_offset=5)
    case("wide_coords_L8_m1", 1, 8, 4, True, big=True)
    merged_case()
