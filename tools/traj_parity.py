"""Long-trajectory parity: CUDA path vs the CPU oracle on the same perturbed start (SURVEY §8c/§8d protocol, gating
fixture final_scale=0.02).  python tools/traj_parity.py [L] [n] [B]   -> prints final C-alpha rel-L2 (gate 1e-4)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import str2str_oracle as O  # noqa: E402
from str2str_b200 import synthetic  # noqa: E402
from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA  # noqa: E402
from str2str_b200.rigid import Rigid  # noqa: E402
from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig  # noqa: E402
from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.set_num_threads(os.cpu_count() or 8)
params = synthetic.make_state_dict(0, 0.02)
feats = synthetic.make_features(B, L, seed=7)
q, x = synthetic.make_backbone(L, 7)
g = torch.Generator().manual_seed(123)
ax, u, z = torch.randn(B, L, 3, generator=g), torch.rand(B, L, generator=g), torch.randn(B, L, 3, generator=g)
rt = O.forward_marginal(O.quat_to_rotmat(q[None].repeat(B, 1, 1)), x[None].repeat(B, 1, 1), 0.5 * torch.ones(B), feats["residue_mask"], ax, u, z)
t0 = time.time()
with torch.no_grad():
    fin_ref, _, _ = O.forward_backward(params, feats, rt, 0.5, 2 * n)
t_cpu = time.time() - t0

net = DenoisingNet(EmbeddingModule(32, 256, 128), TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64),
                   pair_kernels=1, node_gemm=1)
net.load_state_dict(params, strict=True)
net = net.cuda().eval()
diffuser = FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir="/tmp/str2str_b200_cache"), min_t=1e-2)
smp = ForwardBackwardSampler(net, diffuser, InferenceConfig(num_timesteps=2 * n, min_t=0.01), use_cuda_graph=True)
one = synthetic.make_features(1, L, seed=7)
r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
_, fin, _ = smp.forward_backward({k: v.cuda() for k, v in one.items()}, r0, 0.5, rigids_t=rt.cuda(), return_rigids=True)
ca, ca_ref = fin.cpu()[..., 4:].double(), fin_ref[..., 4:].double()
r = float((ca - ca_ref).norm() / ca_ref.norm())
print(f"traj parity L={L} n={n} B={B} [{os.environ.get('S2S_TAG', 'default')}]: C-alpha rel-L2 {r:.3e}  max|d| {float((ca - ca_ref).abs().max()):.3e} A  "
      f"(oracle {t_cpu:.1f} s on CPU)")
