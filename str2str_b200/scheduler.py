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
quaternion -> rotation vector -> quaternion re-encoding; tests/test_gpu_parity.py), and is checked against the CPU oracle's
per-delta trajectories directly (tests/test_gpu_production.py).

Control is host-side but vectorised: the phase of every row is a deterministic function of the queue, so one iteration costs a
few numpy operations on [slots]-sized arrays, three small pinned-host -> device copies (schedule-table row index, flags, step
size) and three device gathers; there is no per-slot Python work and no device -> host traffic.  The SDE sampler
(probability_flow = False) draws each row's noise keyed by (seed, trajectory id, iteration) — s2s_rng_fill_rows — i.e. the same
numbers `ForwardBackwardSampler` draws for that decoy when it runs the delta on its own.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .rigid import Rigid
from .sampler import ForwardBackwardSampler
from .score.frame import schedule_rows


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
            rigids_t: Optional[Dict[float, torch.Tensor]] = None, return_rigids: bool = False, seed: Optional[int] = None,
            first_decoy: int = 0):
        """`work` = [(delta, n_replica), ...] for ONE protein (batch size 1).  Returns {delta: atom37 [n_replica, L, 37, 3]}
        (numpy), plus {delta: tensor_7 [n_replica, L, 7]} with `return_rigids`.  `rigids_t[delta]` ([n_replica, L, 7]) optionally
        replaces the internal perturbation (parity tests).  Trajectory (delta #d, replica r) is job-wide decoy
        `first_decoy + sum(n_replica of earlier deltas) + r`: with `seed`, its perturbation and SDE noise are keyed by that id
        (the numbers ForwardBackwardSampler.sample(..., seed, first_decoy=<same id>) draws)."""
        from . import _lib

        smp = self.sampler
        cfg, net, diffuser = smp.cfg, smp.net, smp.diffuser
        sde = not cfg.probability_flow
        if sde and seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)))  # the noise of a continuous batch is always keyed by trajectory
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
        prime = 1 if cfg.self_conditioning else 0

        # ---- the work queue: per trajectory its schedule-table segment, step count, step size, global decoy id ----
        tabs, seg_base, seg_n = [], {}, {}
        starts, t_base, t_n, t_dt, t_delta_idx = [], [], [], [], []
        offs = 0
        decoy0 = first_decoy
        for di, (delta, n_rep) in enumerate(work):
            delta = float(delta)
            T = delta if delta > 0 else 1.0
            if delta not in seg_base:
                n = int(float(cfg.num_timesteps) * T)
                ts = np.linspace(cfg.min_t, T, n)[::-1]
                t32 = torch.as_tensor(ts.copy(), dtype=torch.float64).to(torch.float32)  # t * ones(B) is fp32 in the reference
                rows, _ = schedule_rows(diffuser.trans_diffuser, diffuser.rot_diffuser, t32)
                seg_base[delta], seg_n[delta] = offs, n
                tabs.append(rows)
                offs += n
            if rigids_t is not None and delta in rigids_t:
                st = rigids_t[delta].to(dev, torch.float32).contiguous()
            else:
                r0 = Rigid.from_tensor_4x4(gt.repeat(n_rep, *(1,) * (gt.ndim - 1)))
                if delta > 0:
                    st = diffuser.forward_marginal(r0, delta * torch.ones(n_rep), diffuse_mask=s["rmask64"][:1].expand(n_rep, L),
                                                   as_tensor_7=True, seed=seed, first_decoy=decoy0)["rigids_t"].to(dev, torch.float32).contiguous()
                else:
                    st = diffuser.sample_prior(r0.shape, dev, as_tensor_7=True, seed=seed, first_decoy=decoy0)["rigids_t"].to(dev, torch.float32).contiguous()
            assert st.shape == (n_rep, L, 7), (st.shape, (n_rep, L, 7))
            starts.append(st)
            t_base += [seg_base[delta]] * n_rep
            t_n += [seg_n[delta]] * n_rep
            t_dt += [1.0 / seg_n[delta]] * n_rep
            t_delta_idx += [di] * n_rep
            decoy0 += n_rep
        N = len(t_base)
        t_base, t_n, t_dt = np.asarray(t_base, np.int64), np.asarray(t_n, np.int64), np.asarray(t_dt, np.float64)
        idle_rows, _ = schedule_rows(diffuser.trans_diffuser, diffuser.rot_diffuser, torch.tensor([cfg.min_t], dtype=torch.float32))
        sched_tab = torch.cat(tabs + [idle_rows], 0).to(dev)          # last row: any valid time for idle rows
        idle_idx = sched_tab.shape[0] - 1
        start_all = torch.cat(starts, 0)                              # [N, L, 7]
        out_rig = torch.empty(N, L, 7, device=dev)
        out_psi = torch.empty(N, L, 2, device=dev)

        # ---- static buffers of the captured iteration ----
        f32 = dict(device=dev, dtype=torch.float32)
        state, sc = torch.zeros(S, L, 7, **f32), torch.zeros(S, L, 3, **f32)
        state[..., 0] = 1.0                                     # idle rows: identity frames
        out7, psi = torch.empty(S, L, 7, **f32), torch.empty(S, L, 2, **f32)
        t_cur, sched_cur = torch.empty(S, **f32), torch.empty(S, 8, **f32)
        sched_d = torch.empty(S, 2, device=dev, dtype=torch.float64)
        diffuse_cur = torch.zeros_like(s["diffuse"])
        rot_n = torch.zeros(S, L, 3, **f32) if sde else None
        tr_n = torch.zeros(S, L, 3, **f32) if sde else None
        # pinned staging, double-buffered so that the host can prepare iteration k + 1 while the copies of k are in flight
        pin = lambda *shape, dtype: [torch.empty(*shape, dtype=dtype).pin_memory() for _ in range(2)]
        h_idx, h_flag, h_sd = pin(S, dtype=torch.int64), pin(S, 1, dtype=torch.float32), pin(S, 2, dtype=torch.float64)
        h_decoy, h_stream = pin(S, dtype=torch.int64), pin(2, S, dtype=torch.int32)
        d_idx, d_flag = torch.empty(S, dtype=torch.int64, device=dev), torch.empty(S, 1, **f32)
        d_decoy, d_stream = torch.empty(S, dtype=torch.int64, device=dev), torch.empty(2, S, dtype=torch.int32, device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        used = [False, False]
        lib = _lib.load()

        def step_body():
            eng.net_forward(state, sc, t_cur, s["ridx"], s["rmask"], s["fixed"], s["gt_psi"], out7, psi)
            if cfg.self_conditioning:
                sc.copy_(out7[..., 4:])
            diffuser.score_and_reverse(out7, state, s["rmask"], diffuse_cur, sched_cur, sched_d, state, noise_scale=cfg.noise_scale,
                                       probability_flow=cfg.probability_flow, rot_noise=rot_n, trans_noise=tr_n)

        # ---- slot bookkeeping on the host, vectorised ----
        slot_traj = np.full(S, -1, np.int64)     # trajectory in each row (-1: idle)
        slot_phase = np.zeros(S, np.int64)
        nxt = 0

        def refill(rows_np):
            """Put the next queued trajectories into the given rows (one gather / scatter on the device)."""
            nonlocal nxt
            k = min(len(rows_np), N - nxt)
            slot_traj[rows_np] = -1
            if k > 0:
                rows_k = rows_np[:k]
                ids = np.arange(nxt, nxt + k)
                slot_traj[rows_k], slot_phase[rows_k] = ids, 0
                nxt += k
                r_dev, i_dev = torch.as_tensor(rows_k, device=dev), torch.as_tensor(ids, device=dev)
                state.index_copy_(0, r_dev, start_all.index_select(0, i_dev))
                sc.index_fill_(0, r_dev, 0.0)

        def load_inputs(it):
            b = it & 1
            if used[b]:
                copied[b].synchronize()                     # the copies issued two iterations ago have left this pinned set
            live = slot_traj >= 0
            tr = np.where(live, slot_traj, 0)
            step = np.maximum(0, slot_phase - prime)          # index into the trajectory's schedule segment (priming runs at ts[0])
            last = live & (slot_phase == t_n[tr] + prime - 1)
            priming = live & (slot_phase < prime)
            h_idx[b].numpy()[:] = np.where(live, t_base[tr] + step, idle_idx)
            h_flag[b].numpy()[:, 0] = (live & ~priming & ~last).astype(np.float32)
            sd = np.where(live, t_dt[tr], 1.0)
            h_sd[b].numpy()[:, 0], h_sd[b].numpy()[:, 1] = sd, np.sqrt(sd)
            d_idx.copy_(h_idx[b], non_blocking=True); d_flag.copy_(h_flag[b], non_blocking=True); sched_d.copy_(h_sd[b], non_blocking=True)
            if sde:
                h_decoy[b].numpy()[:] = np.where(live & ~priming & ~last, first_decoy + tr, -1)   # rows that take a noisy step
                h_stream[b].numpy()[0], h_stream[b].numpy()[1] = 16 + 2 * step, 17 + 2 * step
                d_decoy.copy_(h_decoy[b], non_blocking=True); d_stream.copy_(h_stream[b], non_blocking=True)
            copied[b].record()
            used[b] = True
            torch.index_select(sched_tab, 0, d_idx, out=sched_cur)
            t_cur.copy_(sched_cur[:, 0])
            torch.mul(s["diffuse"], d_flag, out=diffuse_cur)
            if sde:
                st_ = _lib.stream(dev)
                _lib.check(lib.s2s_rng_fill_rows(_lib.ptr(rot_n), S, L * 3, int(seed) & (2 ** 64 - 1), _lib.ptr_i64(d_decoy),
                                                 _lib.ptr(d_stream[0], torch.int32), 0, st_))
                _lib.check(lib.s2s_rng_fill_rows(_lib.ptr(tr_n), S, L * 3, int(seed) & (2 ** 64 - 1), _lib.ptr_i64(d_decoy),
                                                 _lib.ptr(d_stream[1], torch.int32), 0, st_))
            return live, last

        graph = None
        self.iterations = self.row_iterations = 0
        with torch.no_grad():
            refill(np.arange(S))
            if smp.use_cuda_graph:
                # warm-up outside capture on the real first inputs (restored afterwards), then capture one iteration
                load_inputs(0)
                snap = (state.clone(), sc.clone())
                step_body()
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step_body()
                state.copy_(snap[0]); sc.copy_(snap[1])
            it = 0
            while (slot_traj >= 0).any():
                live, last = load_inputs(it)
                if graph is not None:
                    graph.replay()
                else:
                    step_body()
                it += 1
                self.iterations += 1
                self.row_iterations += int(live.sum())
                done_rows = np.nonzero(last)[0]
                if len(done_rows):  # harvest the finished rows (the network output itself is the result, :304-305), refill them
                    r_dev = torch.as_tensor(done_rows, device=dev)
                    i_dev = torch.as_tensor(slot_traj[done_rows], device=dev)
                    out_rig.index_copy_(0, i_dev, out7.index_select(0, r_dev))
                    out_psi.index_copy_(0, i_dev, psi.index_select(0, r_dev))
                    refill(done_rows)
                slot_phase[live & ~last] += 1
            atom37, rigs = {}, {}
            lo = 0
            for delta, n in work:
                d = float(delta)
                parts = []
                for a0 in range(lo, lo + n, S):
                    a1 = min(lo + n, a0 + S)
                    aat = s["aatype"][: a1 - a0] if s["aatype"] is not None else None
                    a37, _ = eng.backbone_atoms(out_rig[a0:a1].contiguous(), out_psi[a0:a1].contiguous(), aat, want_atom14=False)
                    parts.append(a37)
                atom37[d] = torch.cat(parts, 0).cpu().numpy()
                rigs[d] = out_rig[lo:lo + n].clone()
                lo += n
        if return_rigids:
            return atom37, rigs
        return atom37
