#!/usr/bin/env python
"""Benchmark of the Str2Str forward-backward denoising hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4|cfg5]
                  [--decoys B] [--length L] [--denoise-steps n] [--no-extras]

A "step" is one full forward_backward over one batch of decoys: perturbation, self-conditioning priming
forward, n denoising iterations (score network + fused SE(3) step), backbone build.  Default workload is
BASELINE.json configs[1]: 256 residues, 100 denoise steps, 64 decoys per GPU.  Prints ONE JSON line (rank 0).

  value  : conformations/sec with inputs resident in HBM (device-timed, max over ranks)
  e2e    : same metric through the public API with HOST inputs: pinned-host feature dict copied H2D and the
           atom37 result copied D2H inside the timed region, every step
  roofline / cpu_baseline: see DESIGN.md "Measurement"
  extra  : (N = 1 only) sub-measurements outside the headline: BASELINE cfg 1 (L=64, B=1, n=10) and a ragged length
           (L=250) on the same tcgen05 path, cfg 4 (L=512, n=200, B=64), and `ref_gpu`: the UNMODIFIED reference
           (.cuda(), fp32, allow_tf32=False) on a bounded sample of cfg 2 — the north star's ">= 10x the reference
           single-GPU PyTorch path" denominator.
  --workload cfg5 : BASELINE configs[4], 256 decoys as 64 each of L = 64 / 128 / 256 / 384, 100 denoise steps, distributed
           over the ranks by cost (sampler.plan_mixed_lengths), un-padded batches per length, one gather of the coordinates.

`--impl reference` times the UNMODIFIED reference (staged by tools/stage_reference.py under the git-ignored baseline/_ref/;
the oracle port if that is absent) on all host cores: BASELINE cfg 1 in full, and for the benchmarked config a bounded
sample (one decoy, as many of the n + 1 forwards as fit the time budget) extrapolated to one conformation.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from str2str_b200 import synthetic  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--decoys", type=int, default=64, help="decoys per GPU (BASELINE cfg2: 64)")
    ap.add_argument("--length", type=int, default=256)
    ap.add_argument("--denoise-steps", type=int, default=100)
    ap.add_argument("--pair-kernels", type=int, default=1)
    ap.add_argument("--node-gemm", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-measurements of the `extra` block")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2: --length/--decoys/--denoise-steps as given (defaults = BASELINE cfg 2); cfg4: L=512, n=200, B=64; cfg5: mixed lengths")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="wall seconds for the whole --impl reference run")
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    if a.workload == "cfg4":
        a.length, a.denoise_steps, a.decoys = 512, 200, 64
    return a


def config_of(a):
    """The workload description, identical in both arms (ours / reference)."""
    L, n, B = a.length, a.denoise_steps, a.decoys
    if a.workload == "cfg5":
        return {"workload": "mixed-length batch {64,128,256,384} residues, 100 denoise steps, 256 decoys (BASELINE cfg5)",
                "L": [64, 128, 256, 384], "denoise_steps": 100, "decoys_total": 256,
                "l2_policy": "inputs larger than L2 for L >= 128 (pair tensor >= 0.27 GB per pass); L = 64 batches are L2-resident"}
    tag = " (BASELINE cfg2)" if (L, n, B) == (256, 100, 64) else " (BASELINE cfg4)" if (L, n, B) == (512, 200, 64) else ""
    return {"workload": f"{L}-residue chain, {n} denoise steps, {B} decoys per GPU{tag}", "L": L, "denoise_steps": n,
            "decoys_per_gpu": B, "network_forwards_per_step": n + 1,
            "l2_policy": f"inputs larger than L2 (pair tensor {B * L * L * 256 / 1e9:.2f} GB bf16 per pass)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_batch(L: int, seed: int = 7):
    """Single-protein feature dict on the HOST (pinned), as the reference datamodule would hand it over."""
    feats = synthetic.make_features(1, L, seed=seed)
    q, x = synthetic.make_backbone(L, seed=seed)
    return feats, q, x


def gt_frames_4x4(q, x):
    """rigidgroups_gt_frames[..., 0, :, :]-style 4x4 backbone frames from (quat, trans)."""
    a, b, c, d = q.unbind(-1)
    R = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
                     2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
                     2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1).reshape(-1, 3, 3)
    M = torch.zeros(q.shape[0], 4, 4)
    M[:, :3, :3] = R
    M[:, :3, 3] = x
    M[:, 3, 3] = 1
    return M


def algorithmic(B, L):
    """Per-launch ALGORITHMIC work of the candidate dominant kernels = SURVEY.md §8(d)'s per-unit figures x the units one launch
    processes (DESIGN.md §6 states them): name -> (bound, algorithmic FLOPs or bytes, algorithmic bytes, extra fields).
      EdgeTransition: F_ET = 2 (2 * 384^2 + 384 * 128) = 688 128 FLOP per pair row — the reference's three 384-wide layers.  The kernel
        executes 655 360 per row (n'_i terms folded into per-residue vectors); `executed_flops` carries that count.
      IPA pair kernel: bytes_IPA = B [L^2 c_z s_z + 2 L c_s 4 + L 7 4 + L 4] + parameters once, with s_z = 2 (bf16 pair tensor).  The
        kernel as built also reads the logits and writes the attention weights (2 B H L^2 4 bytes: the q.k^T and P.v GEMMs are separate
        launches); `moved_bytes` is that larger, actually necessary traffic of this kernel.
      edge embedder: 2 * 2 * 128^2 FLOP per pair row (layer 1 table-ised, as SURVEY §8d allows); writes the pair tensor once."""
    rows = B * L * L
    et_flops = rows * 2 * (2 * 384 * 384 + 384 * 128) + 2 * B * L * 256 * 128
    et_exec = rows * 2 * (256 * 384 + 384 * 384 + 640 * 128)
    et_bytes = rows * 128 * 2 * 2                                           # z read + z' write, bf16
    ipa_alg = B * (L * L * 128 * 2 + 2 * L * 256 * 4 + L * 7 * 4 + L * 4) + 2445264 * 4
    ipa_moved = rows * 128 * 2 + 2 * B * 8 * L * L * 4 + B * L * 256 * 4     # z (bf16) + logits in / weights out + o_pair
    ee_flops = rows * 2 * 2 * 128 * 128
    return dict(edge_transition=("tensor", et_flops, et_bytes, {"executed_flops": et_exec}),
                ipa_pair_attention=("hbm", ipa_alg, ipa_alg, {"moved_bytes": ipa_moved}),
                edge_embed=("tensor", ee_flops, rows * 128 * 2, {}))


NCU_CAPTURE = "r02h_ncu_full_summary.csv"  # `ncu --set full --clock-control none` of tools/profile_forward.py 64 256 1, condensed by tools/ncu_summary.py


def ncu_traffic(B, L):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture at cfg2
    (profiles/NCU_CAPTURE): the average over a kernel's launches, summed over the kernels one profiled stage consists of
    (the edge embedder = table MLP + expansion).  Empty for any other shape."""
    if (B, L) != (64, 256):
        return {}
    import csv

    stages = {"edge_transition_pair_kernel": "edge_transition", "edge_transition_tc3_kernel": "edge_transition", "ipa_pair_tc_kernel": "ipa_pair_attention",
              "edge_embed_pipe_kernel<2>": "edge_embed", "edge_embed_expand_kernel": "edge_embed", "embed_table_setup_kernel": "edge_embed"}
    path = os.path.join(ROOT, "profiles", NCU_CAPTURE)
    if not os.path.exists(path):
        return {}
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ir = next(i for i, h in enumerate(hdr) if h.startswith("dram_read"))
    iw = next(i for i, h in enumerate(hdr) if h.startswith("dram_write"))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    ur, uw = hdr[ir].split("[")[1].rstrip("]"), hdr[iw].split("[")[1].rstrip("]")
    per_kernel = {}
    for r in rows[1:]:
        kern = next((k for k in stages if r[0].startswith(k)), None)
        if kern:
            per_kernel.setdefault(kern, []).append(float(r[ir]) * scale[ur] + float(r[iw]) * scale[uw])
    out = {}
    for kern, vals in per_kernel.items():
        out[stages[kern]] = out.get(stages[kern], 0.0) + sum(vals) / len(vals)
    return out


def roofline_records(per, B, L):
    """per = {kernel name: (total device ms, launches)} of one eager step -> (roofline object of the dominant kernel, per-kernel records)."""
    hbm, tf, src = peaks()
    alg = algorithmic(B, L)
    traffic = ncu_traffic(B, L)
    extra, roof = {}, None
    total_prof = sum(v[0] for v in per.values())
    for name, (tot, cnt) in per.items():
        avg = tot / cnt
        rec = {"launches": cnt, "avg_ms": round(avg, 4), "share_of_profiled": round(tot / total_prof, 3)}
        if name in alg:
            bound, work, byts, more = alg[name]
            if bound == "tensor":
                ach = work / (avg * 1e-3) / 1e12
                rec.update(bound="tensor", achieved=round(ach, 1), peak=tf, unit="TFLOP/s", frac=round(ach / tf, 3))
                if "executed_flops" in more:
                    rec["executed_tflops"] = round(more["executed_flops"] / (avg * 1e-3) / 1e12, 1)
            else:
                ach = work / (avg * 1e-3) / 1e9
                rec.update(bound="hbm", achieved=round(ach, 1), peak=hbm, unit="GB/s", frac=round(ach / hbm, 3))
                if "moved_bytes" in more:
                    rec["moved_gbs"] = round(more["moved_bytes"] / (avg * 1e-3) / 1e9, 1)
                    rec["moved_frac"] = round(more["moved_bytes"] / (avg * 1e-3) / 1e9 / hbm, 3)
            rec["algorithmic_bytes"] = byts
            rec["traffic"] = traffic.get(name)
        extra[name] = rec
    dom = max((k for k in per if k in alg), key=lambda k: per[k][0], default=None)
    if dom:
        roof = {k: extra[dom][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
        roof.update({k: extra[dom][k] for k in ("executed_tflops", "moved_gbs", "moved_frac") if k in extra[dom]})
        roof.update(kernel=dom, algorithmic="SURVEY.md 8(d) per-unit figure x units per launch (bench.py: algorithmic())",
                    traffic=traffic.get(dom), traffic_source="ncu --set full, profiles/" + NCU_CAPTURE + " (dram read + write per launch)",
                    peak_source=src, timing="CUDA events around each launch, one eager step")
    return roof, extra


def build_sampler(a, dev, n=None, B=None):
    from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig
    from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser

    net = DenoisingNet(EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128),
                       TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64),
                       pair_kernels=a.pair_kernels, node_gemm=a.node_gemm)
    net.load_state_dict(synthetic.make_state_dict(seed=0, final_scale=0.02), strict=True)
    net = net.to(dev).eval()
    diffuser = FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir="/tmp/str2str_b200_cache"), min_t=1e-2)
    n, B = n or a.denoise_steps, B or a.decoys
    cfg = InferenceConfig(num_timesteps=2 * n, min_t=0.01, replica_per_batch=B, n_replica=B)
    return net, diffuser, ForwardBackwardSampler(net, diffuser, cfg, use_cuda_graph=True)


def host_inputs(L):
    """Pinned host feature dict of one protein, as the reference datamodule hands it over."""
    feats, q, x = host_batch(L)
    host = {k: v.pin_memory() for k, v in feats.items()}
    host["rigidgroups_gt_frames"] = gt_frames_4x4(q, x)[None, :, None].repeat(1, 1, 8, 1, 1).contiguous().pin_memory()
    return host


def dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def time_workload(smp, host, B, dev, steps, warmup, seed, first_decoy, e2e):
    """Device milliseconds of `steps` forward_backward calls over B decoys (CUDA events on the current stream, after `warmup`
    untimed calls).  e2e: host feature dict H2D and atom37 D2H inside every timed step."""
    from str2str_b200.rigid import Rigid

    def once():
        if e2e:
            batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        else:
            batch = once.resident
        r0 = Rigid.from_tensor_4x4(batch["rigidgroups_gt_frames"][..., 0, :, :].repeat(B, 1, 1, 1))
        return smp.forward_backward(batch, r0, 0.5, return_numpy=e2e, seed=seed, first_decoy=first_decoy)

    once.resident = {k: v.to(dev) for k, v in host.items()}
    for _ in range(warmup):
        once()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = once()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1), out


def run_ours(a):
    import torch.distributed as dist

    from str2str_b200 import _lib
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig, all_gather_decoys, shard_bounds

    world, rank, local = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.workload == "cfg5":
        run_cfg5(a, world, rank, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    B, L, n = a.decoys, a.length, a.denoise_steps
    net, diffuser, smp = build_sampler(a, dev)
    host = host_inputs(L)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    d2h_bytes = B * L * 37 * 3 * 4
    first = rank * B  # job-wide decoy ids: the draws of a decoy do not depend on the world size

    def step_resident(dev_batch, r0):
        return smp.forward_backward(dev_batch, r0, 0.5, return_numpy=False, seed=a.seed, first_decoy=first)

    def step_e2e():
        dev_batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        gt = dev_batch["rigidgroups_gt_frames"][..., 0, :, :]
        r0 = Rigid.from_tensor_4x4(gt.repeat(B, 1, 1, 1))
        return smp.forward_backward(dev_batch, r0, 0.5, return_numpy=True, seed=a.seed, first_decoy=first)  # includes the D2H copy of atom37

    dev_batch = {k: v.to(dev) for k, v in host.items()}
    r0 = Rigid.from_tensor_4x4(dev_batch["rigidgroups_gt_frames"][..., 0, :, :].repeat(B, 1, 1, 1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k, gather):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            out = fn()
            if gather and world > 1:  # the single collective of the path: final coordinates
                all_gather_decoys(torch.as_tensor(out)[..., :5, :].contiguous(), shard_bounds(B * world, world))
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(a.warmup):
        step_resident(dev_batch, r0)
    clocks = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: step_resident(dev_batch, r0), a.steps, True)
    clk = clocks.stop() if clocks else None
    launches = smp.launches * a.steps
    step_e2e()
    ms_e2e = timed(step_e2e, a.steps, True)

    value = world * B * a.steps / (ms / 1e3)
    e2e = world * B * a.steps / (ms_e2e / 1e3)

    # per-kernel device time of one eager step (events around each launch), for the roofline line
    roof, extra_k = None, {}
    if rank == 0:
        lib = _lib.load()
        eager = ForwardBackwardSampler(net, diffuser, InferenceConfig(num_timesteps=8, min_t=0.01), use_cuda_graph=False)
        lib.s2s_profile_reset()
        lib.s2s_profile_enable(1)
        eager.forward_backward(dev_batch, r0, 0.5, return_numpy=False)
        torch.cuda.synchronize(dev)
        lib.s2s_profile_enable(0)
        per = {}
        for name in ("edge_transition", "ipa_pair_attention", "edge_embed", "gemm", "gemm_tc"):
            tot, cnt = C.c_double(0), C.c_int64(0)
            if lib.s2s_profile_read(name.encode(), C.byref(tot), C.byref(cnt)) == 0 and cnt.value:
                per[name] = (tot.value, cnt.value)
        roof, extra_k = roofline_records(per, B, L)

    extras = None
    if rank == 0 and world == 1 and not a.no_extras:
        extras = run_extras(a, dev, smp, value)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:  # reported at N=1 only (the host cores are shared by the ranks otherwise)
        cpu = cpu_baseline(L, n, a.cpu_budget_s)

    if rank == 0:
        cfg = config_of(a)
        cfg.update(pair_kernels="tcgen05" if a.pair_kernels else "simt", node_gemm="tensor-core" if a.node_gemm else "fp32")
        line = {
            "metric": "conformations/sec (256-res, 100 denoise steps)", "value": round(value, 3), "unit": "conformations/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 node track / bf16 pair track (fp32 accumulate)",
            "data": "synthetic (seeded random-walk backbone, seeded synthetic weights; no checkpoint is reachable offline)",
            "config": cfg,
            "e2e": {"value": round(e2e, 3), "unit": "conformations/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": round(ms_e2e / a.steps, 2)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "kernels": extra_k, "cpu_baseline": cpu, "extra": extras,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_extras(a, dev, smp_main, value_main):
    """Sub-measurements at N = 1 (not part of the headline): other BASELINE configs on the same path and the reference on the GPU."""
    out = {}

    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # an extra must never take the headline down
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.synchronize(dev)

    def sub(L, B, n, steps, warmup):
        from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

        cfg = InferenceConfig(num_timesteps=2 * n, min_t=0.01, replica_per_batch=B, n_replica=B)
        smp = ForwardBackwardSampler(smp_main.net, smp_main.diffuser, cfg, use_cuda_graph=True)  # same weights / native context
        ms, _ = time_workload(smp, host_inputs(L), B, dev, steps, warmup, a.seed, 0, e2e=False)
        ms2, _ = time_workload(smp, host_inputs(L), B, dev, steps, 0, a.seed, 0, e2e=True)
        return {"L": L, "decoys": B, "denoise_steps": n, "steps": steps, "value": round(B * steps / (ms / 1e3), 3),
                "e2e": round(B * steps / (ms2 / 1e3), 3), "unit": "conformations/s", "ms_per_step": round(ms / steps, 2),
                "pair_rows_per_s": round(B * L * L * (n + 1) * steps / (ms / 1e3), 0)}

    if (a.length, a.decoys, a.denoise_steps) == (256, 64, 100):
        guard("cfg1_L64_B1_n10", lambda: sub(64, 1, 10, 5, 3))
        guard("L250_B64_n100", lambda: dict(sub(250, 64, 100, 1, 3), vs_L256_pair_row_rate=None))
        if "value" in out.get("L250_B64_n100", {}):
            r = out["L250_B64_n100"]
            r["vs_L256_pair_row_rate"] = round(r["pair_rows_per_s"] / (value_main * 256 * 256 * 101), 3)
        guard("cfg4_L512_B64_n200", lambda: sub(512, 64, 200, 1, 1))
    guard("ref_gpu", lambda: ref_gpu(a, dev, value_main))
    return out


def ref_gpu(a, dev, value_main):
    """The north star's denominator: the UNMODIFIED reference moved to the GPU with .cuda() (fp32, allow_tf32=False: torch's
    defaults for matmul are left as the reference leaves them), restated loop of tools/ref_loop.py, on a bounded sample of the
    benchmarked config: all decoys of the batch, the priming forward + 4 iterations, extrapolated to n + 1 forwards."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_loop

    if not ref_loop.available():
        return {"unavailable": "reference sources not staged (baseline/_ref missing)"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L, n, B = a.length, a.denoise_steps, a.decoys
    net, diffuser = ref_loop.build_reference(device=dev)
    q, x = synthetic.make_backbone(L, seed=7)
    sync = lambda: torch.cuda.synchronize(dev)
    while B >= 1:
        try:
            feats = {k: v.to(dev) for k, v in synthetic.make_features(B, L, seed=7).items()}
            g = torch.Generator().manual_seed(3)
            rt = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.3 * torch.randn(B, L, 7, generator=g)).float().to(dev)
            ref_loop.forward_backward(net, diffuser, feats, rt, 0.5, 2 * n, max_forwards=2, sync=sync)  # warm-up (cuBLAS handles, allocator)
            _, _, _, times = ref_loop.forward_backward(net, diffuser, feats, rt, 0.5, 2 * n, max_forwards=5, sync=sync)
            break
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            B //= 2
    else:
        return {"unavailable": "out of memory at every batch size"}
    per_fwd = float(np.median(times[1:]))          # iterations (forward + score + reverse); the priming forward has no step
    sec = times[0] + per_fwd * n
    v = B / sec
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    del net, diffuser
    torch.cuda.empty_cache()
    return {"value": round(v, 4), "unit": "conformations/s", "kind": "unmodified reference (baseline/_ref), torch CUDA eager, fp32, allow_tf32=False",
            "decoys": B, "sample": f"{len(times)} of {n + 1} forwards of {B} decoys at L={L} ({per_fwd * 1e3:.0f} ms per iteration), extrapolated",
            "speedup_of_value": round(value_main / v, 1), "peak_mem_gb": round(peak_gb, 1)}


def run_cfg5(a, world, rank, dev):
    """BASELINE configs[4]: 256 decoys as 64 each of L = 64 / 128 / 256 / 384, 100 denoise steps.  Batches are planned over the
    ranks by cost (sampler.plan_mixed_lengths: LPT on B * L^2), every batch runs un-padded at its own length through the same
    captured-graph sampler (one graph per length), then ONE gather of the packed C-alpha coordinates."""
    import torch.distributed as dist

    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import plan_mixed_lengths

    counts = {64: 64, 128: 64, 256: 64, 384: 64}
    n = 100
    plan = plan_mixed_lengths(counts, world, replica_per_batch=a.decoys)
    mine = plan[rank]
    net, diffuser, smp = build_sampler(a, dev, n=n, B=a.decoys)
    hosts = {L: host_inputs(L) for L in counts}
    resident = {L: {k: v.to(dev) for k, v in hosts[L].items()} for L in counts}
    # job-wide decoy ids: decoys of length L are numbered after those of the shorter lengths, batches in plan order
    base = {L: sum(c for l2, c in counts.items() if l2 < L) for L in counts}
    used = {L: 0 for L in counts}
    my_first = []
    for r in range(world):
        for (L, b) in plan[r]:
            if r == rank:
                my_first.append(base[L] + used[L])
            used[L] += b

    busy_marks = []

    def run_all(e2e):
        outs = []
        start_ev = torch.cuda.Event(enable_timing=True)
        start_ev.record()
        busy_marks.append(start_ev)
        for (L, b), fd in zip(mine, my_first):
            batch = {k: v.to(dev, non_blocking=True) for k, v in hosts[L].items()} if e2e else resident[L]
            r0 = Rigid.from_tensor_4x4(batch["rigidgroups_gt_frames"][..., 0, :, :].repeat(b, 1, 1, 1))
            o = smp.forward_backward(batch, r0, 0.5, return_numpy=False, seed=a.seed, first_decoy=fd)
            outs.append(o[:, :, 1, :].reshape(-1, 3))   # C-alpha, packed [b * L, 3]
        packed = torch.cat(outs, 0) if outs else torch.zeros(0, 3, device=dev)
        done_ev = torch.cuda.Event(enable_timing=True)  # this rank's own work ends here; the collective below waits for the slowest rank
        done_ev.record()
        busy_marks.append(done_ev)
        if world > 1:  # one collective: packed coordinates, padded to the largest rank share (cu_seqlens are the plan itself)
            sizes = [sum(b * L for L, b in plan[r]) for r in range(world)]
            buf = torch.zeros(max(sizes), 3, device=dev)
            buf[: packed.shape[0]] = packed
            allb = torch.empty(world * max(sizes), 3, device=dev)
            dist.all_gather_into_tensor(allb, buf)
            packed = allb
        return packed.cpu() if e2e else packed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(1, a.warmup)):
        run_all(False)
    res = {}
    clocks = ClockSampler(dev.index) if rank == 0 else None
    for mode in ("value", "e2e"):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del busy_marks[:]
        e0.record()
        for _ in range(a.steps):
            run_all(mode == "e2e")
        e1.record()
        torch.cuda.synchronize(dev)
        total = torch.tensor([e0.elapsed_time(e1)], device=dev)
        # this rank's own work: from the start of every pass to its last kernel before the collective
        own = torch.tensor([sum(busy_marks[i].elapsed_time(busy_marks[i + 1]) for i in range(0, len(busy_marks), 2))], device=dev)
        barrier()
        mx, bmx, bmn = total.clone(), own.clone(), own.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(bmx, op=dist.ReduceOp.MAX)
            dist.all_reduce(bmn, op=dist.ReduceOp.MIN)
        res[mode] = (float(mx), float(bmx), float(bmn))
    clk = clocks.stop() if clocks else None
    if rank == 0:
        tot = sum(counts.values())
        from str2str_b200.sampler import batch_cost

        ms, busy_max, busy_min = res["value"]
        cost = [sum(batch_cost(L, b) for L, b in plan[r]) for r in range(world)]
        line = {
            "metric": "conformations/sec (mixed lengths, 100 denoise steps)", "value": round(tot * a.steps / (ms / 1e3), 3), "unit": "conformations/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(1, a.warmup), "ms_per_step": round(ms / a.steps, 2), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "fp32 node track / bf16 pair track (fp32 accumulate)", "data": "synthetic",
            "config": dict(config_of(a), plan=[[list(lb) for lb in plan[r]] for r in range(world)]),
            "e2e": {"value": round(tot * a.steps / (res["e2e"][0] / 1e3), 3), "unit": "conformations/s",
                    "h2d_bytes_per_step": int(sum(sum(v.numel() * v.element_size() for v in hosts[L].values()) for L, _ in mine)),
                    "d2h_bytes_per_step": int(sum(b * L for L, b in mine) * 12 * (world if world > 1 else 1))},
            "gpu_launches": int(smp.launches * len(mine) * a.steps), "clocks": clk,
            "imbalance": {"busy_ms_max": round(busy_max / a.steps, 1), "busy_ms_min": round(busy_min / a.steps, 1),
                          "min_over_max": round(busy_min / busy_max, 3), "planned_cost_min_over_max": round(min(cost) / max(cost), 3),
                          "planned_ms_per_rank": [round(c, 1) for c in cost]},
        }
        print(json.dumps(line))


def _cpu_threads():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


def reference_cpu_sample(L, n, B, budget_s, full=False):
    """Seconds per conformation of the UNMODIFIED reference on the host cores: one call of the restated loop (tools/ref_loop.py)
    over B decoys, bounded to the forwards that fit `budget_s` (all n + 1 when `full`)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_loop

    net, diffuser = reference_cpu_sample.cache.setdefault("mods", ref_loop.build_reference(device="cpu"))
    feats = synthetic.make_features(B, L, seed=7)
    q, x = synthetic.make_backbone(L, seed=7)
    g = torch.Generator().manual_seed(3)
    rt = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.3 * torch.randn(B, L, 7, generator=g)).float()
    if full:
        _, _, _, times = ref_loop.forward_backward(net, diffuser, feats, rt, 0.5, 2 * n)
        return sum(times) / B, len(times), times
    _, _, _, probe = ref_loop.forward_backward(net, diffuser, feats, rt, 0.5, 2 * n, max_forwards=2)
    m = int(min(n + 1, max(3, budget_s / max(probe[1], 1e-3))))
    _, _, _, times = ref_loop.forward_backward(net, diffuser, feats, rt, 0.5, 2 * n, max_forwards=m)
    if len(times) == n + 1:
        return sum(times) / B, len(times), times
    it = float(np.median(times[1:]))
    return (times[0] + it * n) / B, len(times), times   # priming forward + n iterations (the last one has no diffusion step: < 2 % of an iteration)


reference_cpu_sample.cache = {}


def cpu_baseline(L: int, n: int, budget_s: float):
    """The reference's CPU path on the host cores (all of them): the UNMODIFIED reference when its sources are staged
    (kind "reference"), else the oracle port (kind "port").  One decoy, a bounded sample of the n + 1 forwards."""
    cores = _cpu_threads()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_loop

    if ref_loop.available():
        sec, m, times = reference_cpu_sample(L, n, 1, budget_s)
        how = "all" if m == n + 1 else f"{m} of"
        return {"value": round(1.0 / sec, 6), "unit": "conformations/s", "cores": cores, "kind": "reference",
                "sample": f"unmodified reference (baseline/_ref), 1 decoy, L={L}: {how} {n + 1} network forwards + diffusion steps "
                          f"({np.median(times[1:]):.3f} s per iteration)" + ("" if m == n + 1 else ", extrapolated to one conformation")}
    return cpu_baseline_port(L, n, 3, cores)


def cpu_baseline_port(L: int, n: int, forwards: int, cores: int):
    """Oracle port on the host cores: `forwards` network forwards + diffusion steps of ONE decoy, extrapolated to the n+1
    forwards / n-1 diffusion steps of a full conformation (used only when the reference sources are not staged)."""
    from oracle import str2str_oracle as O

    params = synthetic.make_state_dict(0, 0.02)
    feats = synthetic.make_features(1, L, seed=7)
    q, x = synthetic.make_backbone(L, seed=7)
    f = dict(feats)
    f["rigids_t"] = torch.cat([q, x], -1)[None]
    f["sc_ca_t"] = torch.zeros(1, L, 3)
    f["t"] = torch.tensor([0.5])
    diffuse = (1 - f["fixed_mask"]) * f["residue_mask"]
    t_fwd, t_step = [], []
    with torch.no_grad():
        O.denoising_net(params, f)  # warm-up
        for _ in range(forwards):
            t0 = time.perf_counter()
            out = O.denoising_net(params, f)
            t_fwd.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            rs, ts = O.diffuser_score(out["rigids"], f["rigids_t"], f["t"], f["residue_mask"])
            f["rigids_t"] = O.diffuser_reverse(f["rigids_t"], rs, ts, f["t"], 1.0 / n, diffuse)
            t_step.append(time.perf_counter() - t0)
            f["sc_ca_t"] = out["rigids"][..., 4:]
    per_conf = float(np.median(t_fwd)) * (n + 1) + float(np.median(t_step)) * (n - 1)
    return {"value": round(1.0 / per_conf, 6), "unit": "conformations/s", "cores": cores, "kind": "port",
            "sample": f"oracle port, 1 decoy, L={L}: {forwards} of {n + 1} network forwards ({np.median(t_fwd):.2f} s each) and diffusion steps "
                      f"({np.median(t_step) * 1e3:.1f} ms each), extrapolated to one conformation"}


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    world, rank, _ = dist_env()
    if rank != 0:
        return
    cores = _cpu_threads()
    L, n, B = (256, 100, 64) if a.workload == "cfg5" else (a.length, a.denoise_steps, a.decoys)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_loop

    t_start = time.perf_counter()
    vals, cfg1 = [], None
    if ref_loop.available():
        kind = "reference"
        sec1, m1, _ = reference_cpu_sample(64, 10, 1, 0, full=True)  # BASELINE cfg 1, in full: 64 residues, 10 denoise steps, batch 1
        cfg1 = {"workload": "single 64-residue synthetic backbone, 10 denoise steps, batch=1 (BASELINE cfg1), run in full",
                "value": round(1.0 / sec1, 4), "unit": "conformations/s", "s_per_conformation": round(sec1, 3), "forwards": m1}
        per_step = a.ref_budget_s / (a.steps + min(a.warmup, 1))
        if a.warmup:
            reference_cpu_sample(L, n, 1, min(per_step, 10.0))
        for _ in range(a.steps):
            sec, m, times = reference_cpu_sample(L, n, 1, per_step)
            vals.append((1.0 / sec, m, float(np.median(times[1:]))))
        v = float(np.median([x[0] for x in vals]))
        m = vals[-1][1]
        sample = (f"unmodified reference (baseline/_ref), per step 1 decoy of {B}, L={L}: {'all' if m == n + 1 else str(m) + ' of'} {n + 1} network "
                  f"forwards + diffusion steps ({vals[-1][2]:.3f} s per iteration)" + ("" if m == n + 1 else ", extrapolated to one conformation"))
    else:
        kind = "port"
        for _ in range(a.steps):
            vals.append((cpu_baseline_port(L, n, 2, cores)["value"], 2, 0.0))
        v = float(np.median([x[0] for x in vals]))
        sample = f"oracle port (reference sources not staged), 1 decoy, L={L}: 2 of {n + 1} forwards per step, extrapolated"
    print(json.dumps({
        "impl": "reference", "metric": "conformations/sec (256-res, 100 denoise steps)", "value": round(v, 6), "unit": "conformations/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * (time.perf_counter() - t_start) / max(1, a.steps), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": dict(config_of(a), pair_kernels="tcgen05" if a.pair_kernels else "simt", node_gemm="tensor-core" if a.node_gemm else "fp32"),
        "cpu_baseline": {"value": round(v, 6), "unit": "conformations/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 6), "unit": "conformations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cfg1": cfg1,
    }))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
