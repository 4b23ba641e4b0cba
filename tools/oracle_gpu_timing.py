"""Context number: the PyTorch restatement of the reference (oracle) run ON THE GPU with stock ATen/cuBLAS fp32
(allow_tf32 off, as the reference leaves it).  This stands in for "the reference single-GPU PyTorch path" of the north
star, which cannot travel to the GPU box.  python tools/oracle_gpu_timing.py [B] [L] [n_forwards]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import str2str_oracle as O
from str2str_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
params = {k: v.cuda() for k, v in synthetic.make_state_dict(0, 0.02).items()}
feats = {k: v.cuda() for k, v in synthetic.make_features(B, L, seed=7).items()}
q, x = synthetic.make_backbone(L, seed=7)
torch.set_default_device("cuda")  # the oracle creates its small constant tensors on the default device
feats["rigids_t"] = torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda()
feats["sc_ca_t"] = x[None].repeat(B, 1, 1).cuda()
feats["t"] = torch.full((B,), 0.5)
diffuse = (1 - feats["fixed_mask"]) * feats["residue_mask"]
with torch.no_grad():
    out = O.denoising_net(params, feats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = O.denoising_net(params, feats)
        rs, ts = O.diffuser_score(out["rigids"], feats["rigids_t"], feats["t"], feats["residue_mask"])
        new = O.diffuser_reverse(feats["rigids_t"], rs, ts, feats["t"], 0.01, diffuse)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
print(f"oracle on GPU (torch fp32): B={B} L={L}: {dt*1e3:.1f} ms per denoise step -> {B / (dt * 101):.3f} conformations/s at 100 steps "
      f"(peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB)")
