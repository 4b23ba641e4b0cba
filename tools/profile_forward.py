"""Run a few score-network forwards at benchmark size (for ncu): python tools/profile_forward.py [B] [L] [n_forwards]"""
import os
import sys

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from str2str_b200 import synthetic  # noqa: E402
from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
net = DenoisingNet(EmbeddingModule(32, 256, 128), TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4,
                                                                  skip_embed_size=64), pair_kernels=1, node_gemm=1)
net.load_state_dict(synthetic.make_state_dict(0, 0.02), strict=True)
net = net.cuda().eval()
feats = {k: v.cuda() for k, v in synthetic.make_features(B, L, seed=7).items()}
q, x = synthetic.make_backbone(L, seed=7)
feats["rigids_t"] = torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda()
feats["sc_ca_t"] = x[None].repeat(B, 1, 1).cuda()
feats["t"] = torch.full((B,), 0.4, device="cuda")
with torch.no_grad():
    for _ in range(n):
        out = net(feats, as_tensor_7=True)
torch.cuda.synchronize()
print("ok", float(out["rigids"].abs().sum()))
