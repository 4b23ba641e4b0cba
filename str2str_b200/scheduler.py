"""Continuous batching of (delta, replica) trajectories: the caller of the hot path, the delta sweep of
`DiffusionLitModule.predict_step` (reference src/models/diffusion_module.py:229-247, 339-367), runs every delta to completion
in batches of `replica_per_batch`, so the last batch of every delta is partly empty (100 replicas = 64 + 36) and trajectories
of different lengths (n = int(num_timesteps * delta) denoising steps) never share a batch.

Here all (delta, replica) trajectories of one protein form ONE work queue feeding a persistent batch of `slots` rows: every row
carries its own time, schedule scalars, step size and phase, a finished row is refilled with the next trajectory on the spot,
and the captured per-iteration CUDA graph of the sampler is replayed until the queue is dry.  Nothing in the kernels changes:
`t`, the schedule rows and `dt` were per-decoy inputs all along (include/str2str_b200.h: s2s_net_forward, s2s_se3_step), and
the reference's per-trajectory control flow maps onto per-row inputs:

  * priming forward (self-conditioning, :294-297): the row runs the iteration with its `diffuse_mask` row set to 0, so the fused
    score + reverse step leaves its frames untouched (frame.py:206-208: new = m * new + (1 - m) * old) while the
    self-conditioning coordinates are taken from the network output;
  * denoising iterations: the row's own (t, sigma bucket, g^2, beta terms, dt);
  * last iteration (t == min_t, :304-305): the network output itself is the result; the row is harvested and refilled.

Decoys do not interact anywhere on the path (the only cross-row reduction, centring, is per decoy), so every
trajectory gets the values it gets from `ForwardBackwardSampler.forward_backward` up to fp32 rounding noise (measured 1.7e-6
relative on the final C-alpha: the priming iteration passes the unchanged frames once through the step kernel's
quaternion -> rotation vector -> quaternion re-encoding; tests/test_gpu_parity.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .rigid import Rigid
from .sampler import ForwardBackwardSampler
from .score.frame import schedule_rows


class Trajectory:
    """One (delta, replica) work item: n loop iterations of the reference closure, preceded by the priming forward."""

    __slots__ = ("delta", "replica", "n", "ts", "dt", "rows", "phase", "n_phases", "prime")

    def __init__(self, delta: float, replica: int, cfg, rows_cache: Dict[float, Tuple[np.ndarray, torch.Tensor]], diffuser):
        T = delta if delta > 0 else 1.0
        self.delta, self.replica = delta, replica
        if delta not in rows_cache:
            n = int(float(cfg.num_timesteps) * T)
            ts = np.linspace(cfg.min_t, T, n)[::-1]
            t32 = torch.as_tensor(ts.copy(), dtype=torch.float64).to(torch.float32)  # t * ones(B) is fp32 in the reference
            rows, _ = schedule_rows(diffuser.trans_diffuser, diffuser.rot_diffuser, t32)
            rows_cache[delta] = (ts, rows)
        self.ts, self.rows = rows_cache[delta]
        self.n = len(self.ts)
        self.dt = 1.0 / self.n
        self.prime = 1 if cfg.self_conditioning else 0
        self.n_phases = self.n + self.prime
        self.phase = 0

    def step_index(self) -> int:
        """Index into ts / rows of the current phase (the priming forward runs at ts[0])."""
        return max(0, self.phase - self.prime)

    def is_priming(self) -> bool:
        return self.phase < self.prime

    def is_last(self) -> bool:
        return self.phase == self.n_phases - 1


def plan_iterations(n_steps: Sequence[int], slots: int, prime: int = 1) -> Tuple[int, int]:
    """(iterations of the continuous schedule, iterations of the reference's per-delta batching) for trajectories of the given
    step counts — the occupancy argument in numbers.  `n_steps[i]` = denoising steps of trajectory i, in queue order; the
    reference runs each distinct step count separately in batches of `slots`."""
    free = [0] * slots                      # iteration at which each slot becomes free (greedy refill = what run() does)
    for n in n_steps:
        k = min(range(slots), key=lambda i: free[i])
        free[k] += n + prime
    by_n: Dict[int, int] = {}
    for n in n_steps:
        by_n[n] = by_n.get(n, 0) + 1
    ref = sum(-(-cnt // slots) * (n + prime) for n, cnt in by_n.items())
    return max(free), ref


class TrajectoryScheduler:
    def __init__(self, sampler: ForwardBackwardSampler, slots: Optional[int] = None):
        self.sampler = sampler
        self.slots = int(slots or sampler.cfg.replica_per_batch)
        self.iterations = 0        # graph replays of the last run
        self.row_iterations = 0    # sum over iterations of occupied rows (occupancy = row_iterations / (iterations * slots))

    def run(self, batch: Dict[str, torch.Tensor], work: Sequence[Tuple[float, int]],
            rigids_t: Optional[Dict[float, torch.Tensor]] = None, return_rigids: bool = False):
        """`work` = [(delta, n_replica), ...] for ONE protein (batch size 1).  Returns {delta: atom37 [n_replica, L, 37, 3]}
        (numpy), plus {delta: tensor_7 [n_replica, L, 7]} with `return_rigids`.  `rigids_t[delta]` ([n_replica, L, 7]) optionally
        replaces the internal perturbation (parity tests)."""
        smp = self.sampler
        cfg, net, diffuser = smp.cfg, smp.net, smp.diffuser
        if not cfg.probability_flow:
            raise NotImplementedError("TrajectoryScheduler runs the probability-flow (ODE) sampler; use ForwardBackwardSampler for SDE runs")
        assert batch["aatype"].shape[0] == 1, "Batch size must be 1 for correct inference."
        gt = batch["rigidgroups_gt_frames"][..., 0, :, :]
        dev = gt.device
        if dev.type != "cuda":
            raise RuntimeError("TrajectoryScheduler needs the batch on a CUDA device (there is no CPU fallback)")
        L = gt.shape[1]
        S = self.slots
        eng = net.native(dev)
        s = smp._static_feats(batch, S, dev)
        eng.reserve(S, L, s["ridx"])

        # ---- the work queue and the perturbed start frames of every trajectory ----
        rows_cache: Dict[float, Tuple[np.ndarray, torch.Tensor]] = {}
        queue: List[Trajectory] = []
        start: Dict[float, torch.Tensor] = {}
        for delta, n_rep in work:
            delta = float(delta)
            if rigids_t is not None and delta in rigids_t:
                start[delta] = rigids_t[delta].to(dev, torch.float32).contiguous()
            else:
                r0 = Rigid.from_tensor_4x4(gt.repeat(n_rep, *(1,) * (gt.ndim - 1)))
                if delta > 0:
                    start[delta] = diffuser.forward_marginal(r0, delta * torch.ones(n_rep), diffuse_mask=s["rmask64"][:1].expand(n_rep, L),
                                                             as_tensor_7=True)["rigids_t"].to(dev, torch.float32).contiguous()
                else:
                    start[delta] = diffuser.sample_prior(r0.shape, dev, as_tensor_7=True)["rigids_t"].to(dev, torch.float32).contiguous()
            assert start[delta].shape == (n_rep, L, 7), (start[delta].shape, (n_rep, L, 7))
            queue.extend(Trajectory(delta, r, cfg, rows_cache, diffuser) for r in range(n_rep))
        out_rig = {float(d): torch.empty(n, L, 7, device=dev) for d, n in work}
        out_psi = {float(d): torch.empty(n, L, 2, device=dev) for d, n in work}

        # ---- static buffers of the captured iteration ----
        f32 = dict(device=dev, dtype=torch.float32)
        state, sc = torch.zeros(S, L, 7, **f32), torch.zeros(S, L, 3, **f32)
        state[..., 0] = 1.0                                     # idle rows: identity frames
        out7, psi = torch.empty(S, L, 7, **f32), torch.empty(S, L, 2, **f32)
        t_cur, sched_cur = torch.empty(S, **f32), torch.empty(S, 8, **f32)
        sched_d = torch.empty(S, 2, device=dev, dtype=torch.float64)
        diffuse_cur = torch.zeros_like(s["diffuse"])
        h_t, h_sched = torch.empty(S, dtype=torch.float32).pin_memory(), torch.empty(S, 8, dtype=torch.float32).pin_memory()
        h_sd, h_flag = torch.empty(S, 2, dtype=torch.float64).pin_memory(), torch.empty(S, 1, dtype=torch.float32).pin_memory()
        flag = torch.empty(S, 1, **f32)
        idle_rows, _ = schedule_rows(diffuser.trans_diffuser, diffuser.rot_diffuser, torch.tensor([cfg.min_t], dtype=torch.float32))

        def step_body():
            eng.net_forward(state, sc, t_cur, s["ridx"], s["rmask"], s["fixed"], s["gt_psi"], out7, psi)
            if cfg.self_conditioning:
                sc.copy_(out7[..., 4:])
            diffuser.score_and_reverse(out7, state, s["rmask"], diffuse_cur, sched_cur, sched_d, state,
                                       noise_scale=cfg.noise_scale, probability_flow=True)

        slots: List[Optional[Trajectory]] = [None] * S
        nxt = 0

        def refill(k):
            nonlocal nxt
            if nxt < len(queue):
                tr = queue[nxt]
                nxt += 1
                slots[k] = tr
                state[k].copy_(start[tr.delta][tr.replica])
                sc[k].zero_()
            else:
                slots[k] = None

        copied = torch.cuda.Event()
        pending = [False]

        def load_inputs():
            if pending[0]:
                copied.synchronize()                            # the previous iteration's copies have left the pinned buffers
            for k, tr in enumerate(slots):
                if tr is None:                                  # idle row: any valid time, frames frozen
                    h_t[k] = float(cfg.min_t); h_sched[k] = idle_rows[0]; h_sd[k, 0] = 1.0; h_sd[k, 1] = 1.0; h_flag[k, 0] = 0.0
                    continue
                i = tr.step_index()
                h_t[k] = float(tr.rows[i, 0]); h_sched[k] = tr.rows[i]
                h_sd[k, 0] = tr.dt; h_sd[k, 1] = np.sqrt(tr.dt)
                h_flag[k, 0] = 0.0 if (tr.is_priming() or tr.is_last()) else 1.0
            t_cur.copy_(h_t, non_blocking=True); sched_cur.copy_(h_sched, non_blocking=True)
            sched_d.copy_(h_sd, non_blocking=True); flag.copy_(h_flag, non_blocking=True)
            copied.record()
            pending[0] = True
            torch.mul(s["diffuse"], flag, out=diffuse_cur)

        graph = None
        self.iterations = self.row_iterations = 0
        with torch.no_grad():
            for k in range(S):
                refill(k)
            if smp.use_cuda_graph:
                # warm-up outside capture on the real first inputs (restored afterwards), then capture one iteration
                load_inputs()
                snap = (state.clone(), sc.clone())
                step_body()
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step_body()
                state.copy_(snap[0]); sc.copy_(snap[1])
            while any(tr is not None for tr in slots):
                load_inputs()
                if graph is not None:
                    graph.replay()
                else:
                    step_body()
                self.iterations += 1
                for k, tr in enumerate(slots):
                    if tr is None:
                        continue
                    self.row_iterations += 1
                    if tr.is_last():
                        out_rig[tr.delta][tr.replica].copy_(out7[k])
                        out_psi[tr.delta][tr.replica].copy_(psi[k])
                        refill(k)
                    else:
                        tr.phase += 1
            atom37 = {}
            for d, n in work:
                d = float(d)
                parts = []
                for lo in range(0, n, S):
                    hi = min(n, lo + S)
                    aat = s["aatype"][: hi - lo] if s["aatype"] is not None else None
                    a37, _ = eng.backbone_atoms(out_rig[d][lo:hi].contiguous(), out_psi[d][lo:hi].contiguous(), aat, want_atom14=False)
                    parts.append(a37)
                atom37[d] = torch.cat(parts, 0).cpu().numpy()
        if return_rigids:
            return atom37, {d: v.clone() for d, v in out_rig.items()}
        return atom37
