"""str2str_b200.metrics against values produced by the unmodified reference functions
(tests/golden/make_golden_metrics.py): rounded metrics equal, intermediate arrays within fp64 reordering noise, integer
clash counts exact.  The same checks run on cuda tensors under -m gpu."""
import glob
import os

import numpy as np
import pytest
import torch

from str2str_b200 import metrics as M

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "metrics_*.npz")))


def _check(path, device):
    g = np.load(path)
    d = {k[7:]: torch.as_tensor(g[k]).to(device) for k in g.files if k.startswith("coords_")}
    w = {"sampled": g["weights_sampled"]}
    got = {"validity": M.validity(d), "validity_k2": M.validity(d, k_exclusion=2), "bonding_validity": M.bonding_validity(d),
           "js_pwd": M.js_pwd(d), "js_pwd_w": M.js_pwd(d, weights=dict(w)), "js_pwd_k1_b20": M.js_pwd(d, n_bins=20, pwd_offset=1),
           "js_rg": M.js_rg(d), "js_rg_w": M.js_rg(d, weights=dict(w))}
    n = 0
    for m, r in got.items():
        for k, v in r.items():
            ref = float(g[f"res_{m}__{k}"])
            assert abs(v - ref) <= 1.0001e-4, (m, k, v, ref)   # both are rounded to 4 decimals: at most one unit apart on a rounding tie
            n += 1
    assert n == 24
    assert np.abs(M.radius_of_gyration(d["target"]).cpu().numpy() - g["arr_rg_target"]).max() < 1e-12
    assert np.abs(M.pairwise_distance_ca(d["target"], k=3).cpu().numpy() - g["arr_pwd3_target"]).max() < 1e-12
    assert np.array_equal(M.steric_clash(d["sampled"]).cpu().numpy(), g["arr_nclash_sampled"])
    return got


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[8:-4] for p in GOLD])
def test_metrics_match_reference(path):
    got = _check(path, "cpu")
    assert got["js_pwd"]["target"] == 0.0 and 0.0 < got["js_pwd"]["sampled"] < 1.0


def test_histogram_columns_is_np_histogram():
    rng = np.random.default_rng(3)
    v = rng.normal(size=(200, 7))
    v[0, 0], v[1, 0] = v[:, 0].min(), v[:, 0].max()
    lo, hi = np.quantile(v, 0.1, axis=0), np.quantile(v, 0.9, axis=0)   # values outside the range are dropped
    lo[3] = hi[3] = 0.25                                                  # degenerate range -> widened by +-0.5
    v[5, 4] = hi[4]                                                       # exactly on the right edge -> last bin
    w = rng.uniform(0.1, 3.0, 200)
    got = M.histogram_columns(torch.as_tensor(v), torch.as_tensor(lo), torch.as_tensor(hi), 13, torch.as_tensor(w)).numpy()
    for d in range(7):
        ref = np.histogram(v[:, d], bins=13, range=(lo[d], hi[d]), weights=w)[0]
        assert np.allclose(got[:, d], ref, rtol=0, atol=1e-12), d


def test_jensenshannon_is_scipy():
    from scipy.spatial import distance

    rng = np.random.default_rng(4)
    p, q = rng.uniform(0, 1, (50, 9)), rng.uniform(0, 1, (50, 9))
    p[:5, 0] = 0.0
    got = M.jensenshannon(torch.as_tensor(p), torch.as_tensor(q)).numpy()
    assert np.allclose(got, distance.jensenshannon(p, q, axis=0), rtol=1e-13, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[8:-4] for p in GOLD])
def test_metrics_on_device(path):
    _check(path, "cuda")
