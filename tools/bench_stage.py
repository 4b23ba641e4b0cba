"""Time individual stages at benchmark size with CUDA events: python tools/bench_stage.py [B] [L]
Env: S2S_PAIR (0/1/2), S2S_IPA (0/1), S2S_WIMG_COPIES."""
import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from str2str_b200 import synthetic
from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
pair = int(os.environ.get("S2S_PAIR", "1"))
net = DenoisingNet(EmbeddingModule(32, 256, 128), TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64), pair_kernels=pair, node_gemm=1)
net.load_state_dict(synthetic.make_state_dict(0, 0.02), strict=True)
net = net.cuda().eval()
eng = net.native("cuda")
if "S2S_IPA" in os.environ:
    eng.set_option("ipa_kernels", int(os.environ["S2S_IPA"]))
f = {k: v.cuda() for k, v in synthetic.make_features(B, L, seed=7).items()}
eng.reserve(B, L, f["residue_idx"])
q, x = synthetic.make_backbone(L, seed=7)
rm = f["residue_mask"].float().contiguous(); fx = f["fixed_mask"].float().contiguous()
t = torch.full((B,), 0.4, device="cuda"); sc = x[None].repeat(B, 1, 1).cuda().contiguous()
node, z = eng.embed(t, f["residue_idx"], fx, sc, rm)
rig = torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda().contiguous()
def timeit(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1)/n:8.3f} ms")
timeit("embed (node+edge)", lambda: eng.embed(t, f["residue_idx"], fx, sc, rm))
timeit("edge_transition stage", lambda: eng.edge_transition(0, node, z, rm))
qn = rig[..., :4].contiguous(); tn = (rig[..., 4:] * 0.1).contiguous()
timeit("ipa stage", lambda: eng.ipa(0, node, z, qn, tn, rm))
gt = f["torsion_angles_sin_cos"][..., 2, :].float().contiguous()
timeit("net_forward", lambda: eng.net_forward(rig, sc, t, f["residue_idx"], rm, fx, gt), n=5)
# per-kernel breakdown of one forward (CUDA events around every launch)
import ctypes as C
lib = eng.lib
lib.s2s_profile_reset(); lib.s2s_profile_enable(1)
for _ in range(3): eng.net_forward(rig, sc, t, f["residue_idx"], rm, fx, gt)
torch.cuda.synchronize(); lib.s2s_profile_enable(0)
buf = C.create_string_buffer(1 << 16)
n = lib.s2s_profile_list(buf, len(buf))
rows = [l.split("\t") for l in buf.value.decode().strip().split("\n")]
rows = sorted(((r[0], float(r[1]) / 3, int(r[2]) // 3) for r in rows), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"per-forward kernel time by name (sum {tot:.2f} ms):")
for name, ms, cnt in rows: print(f"  {ms:8.3f} ms  x{cnt:3d}  {name}")
