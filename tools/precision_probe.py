"""Precision ablation on the CPU oracle: which operands may be rounded to bf16 without breaking the
1e-4 relative C-alpha gate.  Throw-away evidence for DESIGN.md; emulates rounding points of the CUDA design.
usage: python tools/precision_probe.py [L] [n]"""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch, torch.nn.functional as F
from oracle import str2str_oracle as O
from str2str_b200 import synthetic

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fs = float(sys.argv[3]) if len(sys.argv) > 3 else 0.02
torch.set_num_threads(8)
params = synthetic.make_state_dict(0, fs)
feats = synthetic.make_features(1, L, seed=7)
q, x = synthetic.make_backbone(L, 7)
g = torch.Generator().manual_seed(99)
ax, u, z = torch.randn(1, L, 3, generator=g), torch.rand(1, L, generator=g), torch.randn(1, L, 3, generator=g)
rt = O.forward_marginal(O.quat_to_rotmat(q[None]), x[None], torch.tensor([0.5]), feats["residue_mask"], ax, u, z)

bf = lambda t: t.to(torch.bfloat16).float()
def split2(t):
    hi = bf(t); return hi, bf(t - hi)

orig_lin, orig_softmax, orig_einsum = O.lin, torch.softmax, torch.einsum
MODE = {}

def lin(p, name, x):
    w, b = p[name + ".weight"], p[name + ".bias"]
    pair = ("edge_embed" in name) or ("edge_transition" in name and "initial_embed" not in name)
    if pair and MODE.get("pair_bf16"):
        return F.linear(bf(x), bf(w), b)
    if MODE.get("jitter"):
        y = F.linear(x, w, b); return y * (1 + MODE["jitter"] * torch.randn_like(y))
    if not pair and MODE.get("node") == "bf16x3" and w.dim() == 2 and "linear_b" not in name and "down_z" not in name:
        xh, xl = split2(x); wh, wl = split2(w)
        return F.linear(xh, wh) + F.linear(xh, wl) + F.linear(xl, wh) + b
    if not pair and MODE.get("node") == "bf16" and "linear_b" not in name and "down_z" not in name:
        return F.linear(bf(x), bf(w), b)
    if not pair and MODE.get("node") == "tf32" and "linear_b" not in name and "down_z" not in name:
        t32 = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
        return F.linear(t32(x.contiguous()), t32(w), b)
    return orig_lin(p, name, x)

orig_lnorm = O.lnorm
def lnorm(p, name, x):
    y = orig_lnorm(p, name, x)
    if MODE.get("z_bf16") and ("edge_embed" in name or "edge_transition" in name):
        return bf(y)
    return y

orig_ipa = O.ipa
def ipa(p, pre, s, z, quat, trans, mask):
    if not MODE.get("ipa"):
        return orig_ipa(p, pre, s, z, quat, trans, mask)
    # emulate: QK / PV / P*z operands in bf16 (single pass) or split
    mode = MODE["ipa"]
    def einsum(eq, a, b):
        if mode == "bf16": return orig_einsum(eq, bf(a), bf(b))
        if mode == "bf16x2P":  # P split hi+lo, other operand bf16... only used for a-weighted sums
            return orig_einsum(eq, bf(a), bf(b))
        return orig_einsum(eq, a, b)
    torch.einsum = einsum
    try:
        return orig_ipa(p, pre, s, z, quat, trans, mask)
    finally:
        torch.einsum = orig_einsum

O.lin, O.lnorm, O.ipa = lin, lnorm, ipa

def run(**mode):
    MODE.clear(); MODE.update(mode)
    torch.manual_seed(0)
    t0 = time.time()
    fin, psi, a37 = O.forward_backward(params, feats, rt, 0.5, 2 * n)
    return fin[..., 4:], time.time() - t0

base, dt = run()
print(f"L={L} n={n} final_scale={fs}: baseline {dt:.1f}s  |ca| rms {base.pow(2).mean().sqrt():.2f} A")
for name, mode in [
    ("jitter 1e-6", dict(jitter=1e-6)),
    ("jitter 1e-5", dict(jitter=1e-5)),
    ("pair bf16 operands", dict(pair_bf16=True)),
    ("pair bf16 + z bf16 storage", dict(pair_bf16=True, z_bf16=True)),
    ("node bf16x3", dict(node="bf16x3")),
    ("node tf32", dict(node="tf32")),
    ("node bf16", dict(node="bf16")),
    ("ipa einsums bf16", dict(ipa="bf16")),
    ("all: pair bf16 + z bf16 + node bf16x3", dict(pair_bf16=True, z_bf16=True, node="bf16x3")),
    ("all + ipa einsums bf16", dict(pair_bf16=True, z_bf16=True, node="bf16x3", ipa="bf16")),
]:
    ca, dt = run(**mode)
    r = float((ca - base).norm() / base.norm())
    print(f"  {name:45s} rel-L2 {r:.2e}  max|d| {float((ca-base).abs().max()):.2e} A")

# --- which node-side linears tolerate single-pass bf16? -------------------------------------------
import re
def lin_sel(p, name, x):
    w, b = p[name + ".weight"], p[name + ".bias"]
    pat = MODE.get("bf16_names")
    if pat and re.search(pat, name):
        return F.linear(bf(x), bf(w), b)
    return orig_lin(p, name, x)
O.lin = lin_sel
orig_F_linear = F.linear
print("single-pass bf16 on selected node linears:")
for name, pat in [
    ("ipa linear_q/linear_kv", r"ipa_\d\.linear_(q|kv)$"),
    ("ipa q/kv points", r"ipa_\d\.linear_(q|kv)_points$"),
    ("ipa linear_out", r"ipa_\d\.linear_out$"),
    ("ipa linear_b/down_z", r"ipa_\d\.(linear_b|down_z)$"),
    ("transformer ffn+out_proj", r"transformer_\d\.layers\.\d\.(linear1|linear2|self_attn\.out_proj)$"),
    ("node_transition", r"node_transition"),
    ("skip/linear_b/torsion", r"(skip_embed|trunk\.linear_\d|torsion_pred)"),
    ("bb_update", r"bb_update"),
    ("node_embed mlp", r"embedder\.node_embed"),
    ("edge_transition initial_embed", r"initial_embed"),
]:
    ca, dt = run(bf16_names=pat)
    r = float((ca - base).norm() / base.norm())
    print(f"  {name:45s} rel-L2 {r:.2e}  max|d| {float((ca-base).abs().max()):.2e} A")
