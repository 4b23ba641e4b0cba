#!/usr/bin/env python
"""Benchmark of the Str2Str forward-backward denoising hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--decoys B] [--length L] [--denoise-steps n]

A "step" is one full forward_backward over one batch of decoys: perturbation, self-conditioning priming
forward, n denoising iterations (score network + fused SE(3) step), backbone build.  Default workload is
BASELINE.json configs[1]: 256 residues, 100 denoise steps, 64 decoys per GPU.  Prints ONE JSON line (rank 0).

  value  : conformations/sec with inputs resident in HBM (device-timed, max over ranks)
  e2e    : same metric through the public API with HOST inputs: pinned-host feature dict copied H2D and the
           atom37 result copied D2H inside the timed region, every step
  roofline / cpu_baseline: see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU implementation (the oracle port; the reference's own
sources cannot travel to the GPU box) on the host cores for the same config on a bounded sample.
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
    ap.add_argument("--cpu-sample-forwards", type=int, default=3)
    return ap.parse_args()


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


def ncu_traffic(B, L):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/r01c_ncu_full_pair_kernels.csv, made with tools/ncu_summary.py at cfg2).  None for any other shape."""
    if (B, L) != (64, 256):
        return {}
    import csv

    out = {}
    for fn, names in (("r01c_ncu_full_pair_kernels.csv", {"edge_transition_tc": "edge_transition", "ipa_pair_tc_kernel": "ipa_pair_attention",
                                                          "edge_embed_pipe_kernel": "edge_embed"}),):
        path = os.path.join(ROOT, "profiles", fn)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir = next(i for i, h in enumerate(hdr) if h.startswith("dram_read"))
        iw = next(i for i, h in enumerate(hdr) if h.startswith("dram_write"))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        ur, uw = hdr[ir].split("[")[1].rstrip("]"), hdr[iw].split("[")[1].rstrip("]")
        for r in rows[1:]:
            key = next((v for k, v in names.items() if r[0].startswith(k)), None)
            if key:
                out[key] = float(r[ir]) * scale[ur] + float(r[iw]) * scale[uw]
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
                    traffic=traffic.get(dom), traffic_source="ncu --set full, profiles/r01c_ncu_full_pair_kernels.csv (dram read + write per launch)",
                    peak_source=src, timing="CUDA events around each launch, one eager step")
    return roof, extra


def run_ours(a):
    import torch.distributed as dist

    from str2str_b200 import _lib
    from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig, all_gather_decoys, shard_bounds
    from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L, n = a.decoys, a.length, a.denoise_steps

    net = DenoisingNet(EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128),
                       TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64),
                       pair_kernels=a.pair_kernels, node_gemm=a.node_gemm)
    net.load_state_dict(synthetic.make_state_dict(seed=0, final_scale=0.02), strict=True)
    net = net.to(dev).eval()
    diffuser = FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir="/tmp/str2str_b200_cache"), min_t=1e-2)
    cfg = InferenceConfig(num_timesteps=2 * n, min_t=0.01, replica_per_batch=B, n_replica=B)
    smp = ForwardBackwardSampler(net, diffuser, cfg, use_cuda_graph=True)

    feats, q, x = host_batch(L)
    host = {k: v.pin_memory() for k, v in feats.items()}
    host["rigidgroups_gt_frames"] = gt_frames_4x4(q, x)[None, :, None].repeat(1, 1, 8, 1, 1).contiguous().pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    d2h_bytes = B * L * 37 * 3 * 4
    torch.manual_seed(1234 + rank)

    def step_resident(dev_batch, r0):
        return smp.forward_backward(dev_batch, r0, 0.5, return_numpy=False)

    def step_e2e():
        dev_batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        gt = dev_batch["rigidgroups_gt_frames"][..., 0, :, :]
        r0 = Rigid.from_tensor_4x4(gt.repeat(B, 1, 1, 1))
        out = smp.forward_backward(dev_batch, r0, 0.5, return_numpy=True)  # includes the D2H copy of atom37
        return out

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
    roof, extra = None, {}
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
        roof, extra = roofline_records(per, B, L)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:  # reported at N=1 only (the host cores are shared by the ranks otherwise)
        cpu = cpu_baseline(L, n, a.cpu_sample_forwards)

    if rank == 0:
        line = {
            "metric": "conformations/sec (256-res, 100 denoise steps)", "value": round(value, 3), "unit": "conformations/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms / a.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 node track / bf16 pair track (fp32 accumulate)",
            "data": "synthetic (seeded random-walk backbone, seeded synthetic weights; no checkpoint is reachable offline)",
            "config": {"workload": f"{L}-residue chain, {n} denoise steps, {B} decoys per GPU (BASELINE cfg2)" if (L, n, B) == (256, 100, 64)
                       else f"{L}-residue chain, {n} denoise steps, {B} decoys per GPU",
                       "L": L, "denoise_steps": n, "decoys_per_gpu": B, "network_forwards_per_step": n + 1,
                       "l2_policy": "inputs larger than L2 (pair tensor 1.07 GB bf16 per pass)",
                       "pair_kernels": "tcgen05" if a.pair_kernels else "simt", "node_gemm": "tensor-core" if a.node_gemm else "fp32"},
            "e2e": {"value": round(e2e, 3), "unit": "conformations/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": round(ms_e2e / a.steps, 2)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "kernels": extra, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(L: int, n: int, forwards: int):
    """Oracle port on the host cores: `forwards` network forwards + (forwards-1) diffusion steps of ONE decoy,
    extrapolated to the n+1 forwards / n-1 diffusion steps of a full conformation."""
    from oracle import str2str_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
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
            "sample": f"1 decoy, L={L}: {forwards} of {n + 1} network forwards ({np.median(t_fwd):.2f} s each) and diffusion steps "
                      f"({np.median(t_step) * 1e3:.1f} ms each) of the oracle, extrapolated to one conformation"}


def run_reference(a):
    """Reference arm: the reference algorithm's CPU implementation (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L, n, B = a.length, a.denoise_steps, a.decoys
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vals = []
    t0 = time.perf_counter()
    for _ in range(max(1, a.warmup // 3)):
        cpu_baseline(L, n, 1)
    for _ in range(a.steps):
        vals.append(cpu_baseline(L, n, max(1, a.cpu_sample_forwards - 1)))
    v = float(np.median([c["value"] for c in vals]))
    cb = dict(vals[-1])
    cb["value"] = round(v, 6)
    print(json.dumps({
        "impl": "reference", "metric": "conformations/sec (256-res, 100 denoise steps)", "value": round(v, 6), "unit": "conformations/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(1e3 * (time.perf_counter() - t0) / max(1, a.steps), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{L}-residue chain, {n} denoise steps, {B} decoys per GPU (BASELINE cfg2)", "L": L, "denoise_steps": n,
                   "decoys_per_gpu": B},
        "cpu_baseline": cb, "e2e": {"value": round(v, 6), "unit": "conformations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
