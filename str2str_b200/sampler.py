"""Forward-backward conformation sampler (reference src/models/diffusion_module.py:214-369, `predict_step`).

`ForwardBackwardSampler.forward_backward(rigids_0, t_delta)` is the reference's inner closure
(diffusion_module.py:260-334) on the native kernels: one perturbation, one self-conditioning priming forward,
n denoising iterations (network forward + fused score/reverse step), one backbone build, one D2H copy.
The per-iteration body is captured once in a CUDA graph and replayed; step-dependent scalars (t, sigma bucket,
g^2, beta terms) live in small device tables refreshed by two tiny copies per step.

Decoys are independent, so multi-GPU sampling shards the decoy dimension across ranks with no data-path
collective and a single all-gather of the final coordinates (`sample_sharded`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from .net.denoising_ipa import DenoisingNet
from .rigid import Rigid
from .score.frame import FrameDiffuser, schedule_rows

FEATURE_KEYS = ("aatype", "residue_mask", "fixed_mask", "residue_idx", "torsion_angles_sin_cos")


@dataclass
class InferenceConfig:
    """`inference:` block of the reference's configs/model/diffusion.yaml:88-100 (same keys, same defaults)."""
    delta_min: float = 0.25
    delta_max: float = 0.70
    delta_step: float = 0.05
    n_replica: int = 100
    replica_per_batch: int = 64
    num_timesteps: int = 1000
    noise_scale: float = 1.0
    probability_flow: bool = True
    self_conditioning: bool = True
    min_t: float = 1e-2
    output_dir: Optional[str] = None
    backward_only: bool = False


class ForwardBackwardSampler:
    def __init__(self, net: DenoisingNet, diffuser: FrameDiffuser, inference: InferenceConfig, use_cuda_graph: bool = True):
        self.net, self.diffuser, self.cfg = net, diffuser, inference
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, tuple] = {}
        self.launches = 0  # kernels launched by the last forward_backward (library launches; graph replays included)

    # ------------------------------------------------------------------------------------------------
    def _static_feats(self, batch: Dict[str, torch.Tensor], B: int, device):
        """The reference repeats the single-protein features B times (diffusion_module.py:269-272)."""
        rep = lambda v: v.repeat(B, *(1,) * (v.ndim - 1)) if v.shape[0] == 1 else v
        f = {k: rep(batch[k]).to(device) for k in FEATURE_KEYS if k in batch}
        s = {
            "ridx": f["residue_idx"].contiguous(),
            "rmask": f["residue_mask"].to(torch.float32).contiguous(),
            "fixed": f["fixed_mask"].to(torch.float32).contiguous(),
            "gt_psi": f["torsion_angles_sin_cos"][..., 2, :].to(torch.float32).contiguous(),
            "aatype": f["aatype"].contiguous() if "aatype" in f else None,
            "rmask64": f["residue_mask"],
        }
        s["diffuse"] = ((1 - s["fixed"]) * s["rmask"]).contiguous()
        return s

    def forward_backward(self, batch: Dict[str, torch.Tensor], rigids_0: Rigid, t_delta: float, rigids_t: torch.Tensor = None,
                         noises=None, return_numpy: bool = True, return_rigids: bool = False, seed: Optional[int] = None,
                         first_decoy: int = 0):
        """Sample `rigids_0.shape[0]` conformations.  `rigids_t` (tensor_7) optionally replaces the internal
        perturbation (parity tests feed the reference's perturbed state); `noises[k] = (rot, trans)` injects the
        SDE noise of iteration k.  With `seed`, every random draw of decoy b (perturbation and SDE noise) is keyed by the
        job-wide decoy id `first_decoy + b` (FrameDiffuser.decoy_noise), so a decoy's trajectory does not depend on how the job
        is batched or sharded; without it the draws come from torch's device generator like the reference's.
        `t_delta <= 0` (or `cfg.backward_only`) starts from the prior at T = 1 (diffusion_module.py:262-263,280-285)."""
        cfg = self.cfg
        if cfg.backward_only:
            t_delta = 0.0
        T = t_delta if t_delta > 0 else 1.0
        B, L = rigids_0.shape
        dev = rigids_0.device
        n = int(float(cfg.num_timesteps) * T)
        dt = 1.0 / n
        ts = np.linspace(cfg.min_t, T, n)[::-1]
        eng = self.net.native(dev)
        s = self._static_feats(batch, B, dev)
        eng.reserve(B, L, s["ridx"])
        lib0 = int(eng.lib.s2s_launch_count())

        if rigids_t is None:
            if t_delta > 0:
                rigids_t = self.diffuser.forward_marginal(rigids_0, t_delta * torch.ones(B), diffuse_mask=s["rmask64"],
                                                          as_tensor_7=True, seed=seed, first_decoy=first_decoy)["rigids_t"]
            else:
                rigids_t = self.diffuser.sample_prior(rigids_0.shape, dev, as_tensor_7=True, seed=seed, first_decoy=first_decoy)["rigids_t"]
        state = rigids_t.to(dev, torch.float32).contiguous().clone()

        # device tables for all iterations (t identical across decoys, as in the reference)
        t_all = torch.as_tensor(ts.copy(), dtype=torch.float64).to(torch.float32)  # t * ones(B) is fp32 in the reference
        rows, _ = schedule_rows(self.diffuser.trans_diffuser, self.diffuser.rot_diffuser, t_all)
        sched_all = rows[:, None, :].expand(n, B, 8).contiguous().to(dev)
        t_dev_all = t_all[:, None].expand(n, B).contiguous().to(dev)
        sched_d = torch.tensor([[dt, np.sqrt(dt)]] * B, dtype=torch.float64, device=dev)

        sde = not cfg.probability_flow
        # Static buffers + the captured iteration are cached per (shape, mode) and reused by later calls: a call then
        # costs n graph replays and no warm-up iteration / re-capture.  The cache entry dies with the engine workspace
        # it points into (eng.token = (engine uid, generation) changes whenever s2s_reserve re-allocates or the net rebuilds
        # its engine, e.g. after load_state_dict).
        key = (B, L, str(dev), sde, bool(cfg.self_conditioning), float(cfg.noise_scale), bool(self.use_cuda_graph))
        ctx = self._graphs.get(key)
        if ctx is not None and ctx["token"] != eng.token:
            ctx = None  # captured under another engine / workspace allocation: its pointers are dead
        fresh = ctx is None
        if fresh:
            f32 = dict(device=dev, dtype=torch.float32)
            ctx = dict(token=eng.token, graph=None, per_replay=0,
                       state=torch.empty(B, L, 7, **f32), sc=torch.zeros(B, L, 3, **f32), out7=torch.empty(B, L, 7, **f32),
                       psi=torch.empty(B, L, 2, **f32), t_cur=torch.empty(B, **f32), sched_cur=torch.empty(B, 8, **f32),
                       sched_d=torch.empty(B, 2, device=dev, dtype=torch.float64),
                       rot_n=torch.empty(B, L, 3, **f32) if sde else None, tr_n=torch.empty(B, L, 3, **f32) if sde else None,
                       ridx=torch.empty_like(s["ridx"]), rmask=torch.empty_like(s["rmask"]), fixed=torch.empty_like(s["fixed"]),
                       gt_psi=torch.empty_like(s["gt_psi"]), diffuse=torch.empty_like(s["diffuse"]))
            self._graphs[key] = ctx
        for name in ("ridx", "rmask", "fixed", "gt_psi", "diffuse"):
            ctx[name].copy_(s[name])
        ctx["sched_d"].copy_(sched_d)
        ctx["state"].copy_(state)
        ctx["sc"].zero_()
        state, sc, out7, psi, t_cur, sched_cur, sched_d = (ctx[k] for k in ("state", "sc", "out7", "psi", "t_cur", "sched_cur", "sched_d"))
        rot_n, tr_n = ctx["rot_n"], ctx["tr_n"]
        c = ctx

        def net_call():
            eng.net_forward(state, sc, t_cur, c["ridx"], c["rmask"], c["fixed"], c["gt_psi"], out7, psi)

        def step_body():
            net_call()
            if cfg.self_conditioning:
                sc.copy_(out7[..., 4:])
            self.diffuser.score_and_reverse(out7, state, c["rmask"], c["diffuse"], sched_cur, sched_d, state,
                                            noise_scale=cfg.noise_scale, probability_flow=cfg.probability_flow,
                                            rot_noise=rot_n, trans_noise=tr_n)

        with torch.no_grad():
            if cfg.self_conditioning:  # priming forward at ts[0] with zero self-conditioning (:294-297)
                t_cur.copy_(t_dev_all[0])
                net_call()
                sc.copy_(out7[..., 4:])
            if ctx["graph"] is None and self.use_cuda_graph and n > 2:
                # warm-up outside capture (lazy allocations, function attributes), then capture one iteration
                snap = (state.clone(), sc.clone())
                t_cur.copy_(t_dev_all[0]); sched_cur.copy_(sched_all[0])
                if sde:
                    rot_n.normal_(); tr_n.normal_()
                step_body()
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                c0 = int(eng.lib.s2s_launch_count())
                with torch.cuda.graph(graph):
                    step_body()
                ctx["graph"], ctx["per_replay"] = graph, int(eng.lib.s2s_launch_count()) - c0
                lib0 += ctx["per_replay"]  # launches recorded while capturing were not executed
                state.copy_(snap[0]); sc.copy_(snap[1])
            graph, per_replay = (ctx["graph"], ctx["per_replay"]) if n > 2 else (None, 0)
            replays = 0
            for k in range(n):
                t_cur.copy_(t_dev_all[k])
                last = (ts[k] == cfg.min_t)
                if last:  # the final iterate is the network's x0 prediction itself (:304-305)
                    net_call()
                    break
                sched_cur.copy_(sched_all[k])
                if sde:
                    if noises is not None:
                        rot_n.copy_(noises[k][0]); tr_n.copy_(noises[k][1])
                    elif seed is not None:
                        rot_n.copy_(self.diffuser.decoy_noise((B, L, 3), dev, seed, first_decoy, 16 + 2 * k))
                        tr_n.copy_(self.diffuser.decoy_noise((B, L, 3), dev, seed, first_decoy, 17 + 2 * k))
                    else:
                        rot_n.normal_(); tr_n.normal_()
                if graph is not None:
                    graph.replay()
                    replays += 1
                else:
                    step_body()
            pred = out7 if last else state
            atom37, _ = eng.backbone_atoms(pred, psi, s["aatype"], want_atom14=False)
        # each replay executes the captured launches once
        self.launches = int(eng.lib.s2s_launch_count()) - lib0 + replays * per_replay
        result = atom37.detach().cpu().numpy() if return_numpy else atom37
        if return_rigids:
            return result, pred.clone(), psi.clone()
        return result

    # ------------------------------------------------------------------------------------------------
    def sample(self, batch: Dict[str, torch.Tensor], t_delta: float, n_replica: Optional[int] = None, seed: Optional[int] = None,
               first_decoy: int = 0):
        """All replicas of one protein at one delta, chunked by replica_per_batch (:339-352). -> [N,L,37,3] numpy.
        `seed` / `first_decoy`: see forward_backward (replica r is job-wide decoy first_decoy + r)."""
        cfg = self.cfg
        n_replica = cfg.n_replica if n_replica is None else n_replica
        assert batch["aatype"].shape[0] == 1, "Batch size must be 1 for correct inference."
        gt = batch["rigidgroups_gt_frames"][..., 0, :, :]
        outs, done = [], 0
        while done < n_replica:
            bs = min(cfg.replica_per_batch, n_replica - done)
            r0 = Rigid.from_tensor_4x4(gt.repeat(bs, *(1,) * (gt.ndim - 1)))
            outs.append(self.forward_backward(batch, r0, t_delta, seed=seed, first_decoy=first_decoy + done))
            done += bs
        return np.concatenate(outs, axis=0)

    def sample_to_pdb(self, batch: Dict[str, torch.Tensor], t_delta: float, save_to: str, n_replica: Optional[int] = None):
        """`sample` followed by the ensemble write of predict_step (diffusion_module.py:225-232,354-361): all replicas as
        MODEL records of one PDB file, with the protein's own aatype / chain_index / residue_index.  The file is
        byte-identical to the reference's `atom37_to_pdb` (str2str_b200/pdb_writer.py)."""
        from .pdb_writer import atom37_to_pdb

        atom37 = self.sample(batch, t_delta, n_replica)
        extra = {k: batch[k][0].detach().cpu().numpy() for k in ("aatype", "chain_index", "residue_index") if k in batch}
        return atom37_to_pdb(save_to=save_to, atom_positions=atom37, overwrite=True, **extra)

    def sample_sharded(self, batch: Dict[str, torch.Tensor], t_delta: float, n_replica: int, seed: Optional[int] = None):
        """Decoy-sharded sampling under torch.distributed: rank r samples its contiguous share, then ONE
        all_gather of the final atom coordinates.  Works with any backend (nccl on GPUs).
        With `seed`, decoy d draws Philox subsequence d whatever the world size, so the gathered ensemble is the same
        on 1, 2, 4 or 8 ranks (up to the fp32 reordering noise of different batch compositions)."""
        import torch.distributed as dist

        world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
        share = shard_bounds(n_replica, world)
        lo, hi = share[rank], share[rank + 1]
        local = self.sample(batch, t_delta, hi - lo, seed=seed, first_decoy=lo) if hi > lo else np.zeros((0,) + tuple(batch["aatype"].shape[1:]) + (37, 3), np.float32)
        if world == 1:
            return local
        return all_gather_decoys(torch.as_tensor(local), share).numpy()


def batch_cost(L: int, b: int) -> float:
    """Modelled device time (ms) of one 100-step trajectory batch of b decoys of length L on one B200: a per-decoy slope with a
    pair-track term (L^2) and a node-track term, plus a per-batch floor — the dependent-launch latency of 101 iterations, which
    is what makes many small batches expensive.  Fitted (all within 5 %) to bench.py --length L --decoys b, ms per batch:
    (64, 8) 103, (64, 64) 227, (128, 8) 130, (128, 64) 464, (256, 8) 218, (256, 32) 696, (256, 64) 1328, (384, 8) 396,
    (384, 13) 614, (384, 32) 1437.  Only ratios matter to the planner."""
    return b * (2.8885e-4 * L * L - 7.86e-4 * L + 1.177) + max(40.0, 92.0 - 0.12 * L)


def plan_mixed_lengths(counts: Dict[int, int], world: int, replica_per_batch: int = 64, cost=batch_cost):
    """Mixed-length workload (BASELINE cfg 5): every (length, n_decoys) request is cut into k_L nearly equal batches of at most
    `replica_per_batch` decoys and the batches are assigned to ranks by longest-processing-time-first on `cost(L, b)`.  The
    piece counts k_L are chosen to minimise the plan's makespan (every combination for a handful of length classes, one class
    at a time otherwise): a long class is spread over as many ranks as its share of the work, a short one stays in ONE large
    batch instead of an even (and inefficiently small) share per rank.  Returns one list of (L, B) batches per rank.
    Lengths are never padded against each other: each batch runs un-padded at its own L, which is also what defines
    the reference result for this configuration."""
    def pieces(k):
        out = []
        for L, n in counts.items():
            base, rem = divmod(n, k[L])
            out += [(L, base + (1 if i < rem else 0)) for i in range(k[L]) if base + (1 if i < rem else 0) > 0]
        return out

    def lpt(batches):
        load = [0.0] * world
        plan = [[] for _ in range(world)]
        for L, b in sorted(batches, key=lambda lb: -cost(*lb)):
            r = min(range(world), key=lambda q: load[q])
            plan[r].append((L, b))
            load[r] += cost(L, b)
        return max(load), plan

    counts = {L: n for L, n in counts.items() if n > 0}
    kmin = {L: max(1, -(-n // replica_per_batch)) for L, n in counts.items()}
    kmax = {L: min(n, kmin[L] + world - 1) for L, n in counts.items()}
    n_combos = 1
    for L in counts:
        n_combos *= kmax[L] - kmin[L] + 1
    if n_combos <= 20000:  # few length classes (cfg 5 has four): try every combination of piece counts
        import itertools

        best, plan = None, None
        for ks in itertools.product(*[range(kmin[L], kmax[L] + 1) for L in counts]):
            m, p = lpt(pieces(dict(zip(counts, ks))))
            if best is None or m < best * (1.0 - 1e-9):
                best, plan = m, p
        return plan
    k = dict(kmin)  # many classes: raise one piece count at a time while that shortens the makespan
    best, plan = lpt(pieces(k))
    while True:
        cand = None
        for L in counts:
            if k[L] < kmax[L]:
                k2 = dict(k)
                k2[L] += 1
                m, p = lpt(pieces(k2))
                if m < best * (1.0 - 1e-9) and (cand is None or m < cand[0]):
                    cand = (m, p, k2)
        if cand is None:
            return plan
        best, plan, k = cand


def shard_bounds(n: int, world: int):
    """Contiguous, balanced split of n decoys over `world` ranks: bounds[r] .. bounds[r+1]."""
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def all_gather_decoys(local: torch.Tensor, bounds):
    """all_gather of per-rank decoy blocks with (possibly) unequal counts: pad to the max share, gather once, trim."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [bounds[r + 1] - bounds[r] for r in range(world)]
    mx = max(counts)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    pad[: counts[rank]] = local.to(dev)
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad)
    out = out.view((world, mx) + tuple(local.shape[1:]))
    return torch.cat([out[r, : counts[r]] for r in range(world)], 0).cpu()
