"""Golden values for str2str_b200.metrics from the UNMODIFIED reference functions (src/metrics/metrics.py) on seeded ensembles.

    python tests/golden/make_golden_metrics.py

`deeptime` (TICA, used by js_tica only) is not installed: a stub module lets the file import; js_tica is not exercised.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import refshim  # noqa: E402

refshim.install()
for name, attrs in (("deeptime", {}), ("deeptime.decomposition", {"TICA": object})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
from src.metrics import metrics as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def ensemble(rng, B, L, spread, compact=0.0):
    """Persistent random walks (slowly turning direction: few self-contacts) + isotropic noise, optionally pulled to the centre."""
    d = rng.normal(size=(B, 3))
    steps = []
    for _ in range(L):
        d = d + 0.35 * rng.normal(size=(B, 3))
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        steps.append(3.8 * d)
    x = np.cumsum(np.stack(steps, 1), 1)
    x = x - x.mean(1, keepdims=True)
    x = x * (1 - compact) + rng.normal(0, spread, (B, L, 3))
    return x.astype(np.float32)       # sampled coordinates are fp32


def main():
    rng = np.random.default_rng(0)
    cases = {"L24": (24, 40, 30), "L57": (57, 64, 50)}
    for name, (L, Bt, Bs) in cases.items():
        d = {"target": ensemble(rng, Bt, L, 0.05), "sampled": ensemble(rng, Bs, L, 0.6), "compact": ensemble(rng, Bs, L, 0.2, compact=0.7)}
        d["sampled"][0, 3] = d["sampled"][0, 9]                      # a guaranteed clash
        d["sampled"][1, 5] += 9.0                                    # a broken bond
        d["compact"][2] = d["target"][2]                              # values sitting exactly on the reference's min / max edges
        w = {"sampled": rng.uniform(0.5, 2.0, Bs)}
        out = {"validity": R.validity(d), "validity_k2": R.validity(d, k_exclusion=2), "bonding_validity": R.bonding_validity(d),
               "js_pwd": R.js_pwd(d), "js_pwd_w": R.js_pwd(d, weights=dict(w)), "js_pwd_k1_b20": R.js_pwd(d, n_bins=20, pwd_offset=1),
               "js_rg": R.js_rg(d), "js_rg_w": R.js_rg(d, weights=dict(w))}
        flat = {f"{m}__{k}": float(v) for m, r in out.items() for k, v in r.items()}
        extra = {"rg_target": R.radius_of_gyration(d["target"].astype(np.float64)), "pwd3_target": R.pairwise_distance_ca(d["target"].astype(np.float64), k=3),
                 "nclash_sampled": R._steric_clash(d["sampled"].astype(np.float64))}
        np.savez_compressed(os.path.join(OUT, f"metrics_{name}.npz"), weights_sampled=w["sampled"],
                            **{f"coords_{k}": v for k, v in d.items()}, **{f"res_{k}": np.float64(v) for k, v in flat.items()},
                            **{f"arr_{k}": v for k, v in extra.items()})
        print(name, flat)


if __name__ == "__main__":
    main()
