"""Decoy sharding under torch.distributed (NCCL): the gathered ensemble of N ranks against the same job on one rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_parity.py [L] [n_replica] [steps]
With a seed every random draw of decoy d is keyed by (seed, d), so the two ensembles are the same conformations up to the
fp32 reordering noise of different batch compositions."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from str2str_b200 import synthetic  # noqa: E402
from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA  # noqa: E402
from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig  # noqa: E402
from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_rep = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
net = DenoisingNet(EmbeddingModule(32, 256, 128), TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64))
net.load_state_dict(synthetic.make_state_dict(0, 0.02), strict=True)
net = net.to(dev).eval()
diffuser = FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir="/tmp/str2str_b200_cache"), min_t=1e-2)
cfg = InferenceConfig(num_timesteps=2 * steps, min_t=0.01, replica_per_batch=64)
smp = ForwardBackwardSampler(net, diffuser, cfg, use_cuda_graph=True)
feats = synthetic.make_features(1, L, seed=3, random_aatype=True)
q, x = synthetic.make_backbone(L, seed=3)
gt = torch.zeros(1, L, 8, 4, 4)
a, b, c, d = torch.nn.functional.normalize(q, dim=-1).unbind(-1)
gt[0, :, 0, :3, :3] = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c), 2 * (b * c + a * d),
                                   a * a - b * b + c * c - d * d, 2 * (c * d - a * b), 2 * (b * d - a * c), 2 * (c * d + a * b),
                                   a * a - b * b - c * c + d * d], -1).reshape(L, 3, 3)
gt[0, :, 0, :3, 3] = x
gt[0, :, 0, 3, 3] = 1.0
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in dict(feats, rigidgroups_gt_frames=gt).items()}
sharded = smp.sample_sharded(batch, 0.5, n_rep, seed=11)
if rank == 0:
    alone = smp.sample(batch, 0.5, n_rep, seed=11, first_decoy=0)
    ca_s, ca_a = torch.as_tensor(sharded)[:, :, 1], torch.as_tensor(alone)[:, :, 1]
    rel = float((ca_s - ca_a).norm() / ca_a.norm())
    print(f"sharded over {world} rank(s) vs one rank: {n_rep} decoys, L={L}, {steps} steps: C-alpha rel-L2 {rel:.2e}, max|d| {float((ca_s - ca_a).abs().max()):.2e} A, "
          f"ensemble spread {float(ca_a.std(0).mean()):.2f} A")
    assert tuple(ca_s.shape) == (n_rep, L, 3) and rel < 1e-4
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
