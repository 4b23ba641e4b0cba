"""Launch counts and per-kernel profile names of one forward with the node-track chains off / on (B200):
    python tools/chain_launches.py [B] [L]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from str2str_b200 import _lib, synthetic  # noqa: E402
import test_gpu_production as T  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
params = synthetic.make_state_dict(seed=0, final_scale=0.02)
f = synthetic.make_features(B, L, seed=1, n_pad=0, n_fixed=0, random_aatype=True)
q, x = synthetic.make_backbone(L, seed=1)
f["rigids_t"] = torch.cat([q, x], -1)[None].repeat(B, 1, 1).float()
f["sc_ca_t"] = x[None].repeat(B, 1, 1).float()
f["t"] = torch.full((B,), 0.4)
net = T.make_net(params)
lib = _lib.load()
for chain in (0, 1):
    net.set_option("chain", chain)
    with torch.no_grad():
        net(T.cuda(f), as_tensor_7=True)
    torch.cuda.synchronize()
    n0 = lib.s2s_launch_count()
    with torch.no_grad(), T.kernel_log() as kl:
        net(T.cuda(f), as_tensor_7=True)
    print(f"chain={chain}: {lib.s2s_launch_count() - n0} library launches per forward")
    for k, v in sorted(kl.names.items()):
        print(f"    {v:3d}  {k}")
