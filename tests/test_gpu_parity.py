"""GPU parity tests: the CUDA path (through the C ABI, via the host mirror) against the CPU oracle and the
reference's golden vectors.  Run on the B200 box:  python -m pytest tests -m gpu

Tolerances (relative L2 unless stated):
  * fp32 node-track stages ......................... 2e-5
  * pair tensor z (stored in bf16 by design) ....... 6e-3   (bf16 has 8 mantissa bits: 2^-9 = 2e-3 per element)
  * network outputs / trajectories (C-alpha) ....... 1e-4   (BASELINE.json north_star)
"""
import os

import numpy as np
import pytest
import torch

from oracle import str2str_oracle as O
from str2str_b200 import synthetic

pytestmark = pytest.mark.gpu
# (pair_kernels, node_gemm): 0/0 = SIMT pair kernels + exact fp32 node GEMMs (on-device cross-check), 1/1 = tcgen05 everywhere
# (the default; chain lengths that are not a multiple of 32 are padded inside the library).  The production configuration is
# pinned directly against the oracle in tests/test_gpu_production.py.
PAIR_MODES = [(0, 0), (1, 0), (1, 1)]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def load(golden_dir, name):
    return {k: torch.as_tensor(v) for k, v in np.load(os.path.join(golden_dir, name)).items()}


def make_net(params, pair_kernels=1, node_gemm=1):
    if isinstance(pair_kernels, tuple):
        pair_kernels, node_gemm = pair_kernels
    from str2str_b200.net import DenoisingNet, EmbeddingModule, TranslationIPA

    net = DenoisingNet(
        EmbeddingModule(init_embed_size=32, node_embed_size=256, edge_embed_size=128),
        TranslationIPA(c_s=256, c_z=128, coordinate_scaling=0.1, no_ipa_blocks=4, skip_embed_size=64),
        pair_kernels=pair_kernels, node_gemm=node_gemm,
    )
    net.load_state_dict(params, strict=True)
    return net.cuda().eval()


def make_diffuser(tmp="/tmp/str2str_b200_cache"):
    from str2str_b200.score import FrameDiffuser, R3Diffuser, SO3Diffuser

    return FrameDiffuser(R3Diffuser(0.1, 20.0, 0.1), SO3Diffuser(cache_dir=tmp), min_t=1e-2)


def small_feats(g):
    f = synthetic.make_features(2, 12, seed=3, n_pad=2, n_fixed=1, random_aatype=True)
    f.update(rigids_t=g["rigids_t"], sc_ca_t=g["sc_ca_t"], t=g["t"])
    return f


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def test_linear_f32_exact():
    from str2str_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(70, 6, 256), (33, 128, 65), (200, 256, 2688), (64, 64, 16)]:
        A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
        Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
        Cd = torch.empty(M, N, device="cuda")
        _lib.check(lib.s2s_linear_f32(_lib.ptr(Ad), _lib.ptr(Wd), _lib.ptr(bd), _lib.ptr(Cd), M, N, K, 1, _lib.stream()))
        ref = torch.relu(A.double() @ W.double().T + b.double())
        assert rel(Cd, ref) < 2e-6


@pytest.mark.parametrize("pair", PAIR_MODES)
def test_embedder_vs_oracle(golden_dir, params, pair):
    g = load(golden_dir, "net_forward_small.npz")
    f = small_feats(g)
    net = make_net(params, pair)
    fc = cuda(f)
    node, edge = net.embedder(fc["residue_idx"], fc["t"], fc["fixed_mask"], fc["sc_ca_t"])
    node_o, edge_o = O.embedder(params, f["residue_idx"], f["t"], f["fixed_mask"].float(), f["sc_ca_t"])
    assert rel(node, node_o) < 2e-5 and rel(node, g["node_embed"]) < 2e-5
    assert rel(edge, edge_o) < 6e-3 and rel(edge, g["edge_embed"]) < 6e-3
    # a wrong distogram bin or relative-position offset would shift a whole row by O(1)
    assert float((edge.cpu() - edge_o).abs().max()) < 0.08


def test_ipa_block_vs_oracle(golden_dir, params):
    from str2str_b200.rigid import Rigid

    g = load(golden_dir, "net_forward_small.npz")
    f = small_feats(g)
    net = make_net(params, 0, 0)
    nm = f["residue_mask"].float()
    node = g["node_embed"] * nm[..., None]
    edge = (g["edge_embed"] * (nm[..., None] * nm[..., None, :])[..., None]).bfloat16().float()  # what the kernel is fed
    rig = f["rigids_t"].clone()
    rig[..., 4:] *= 0.1
    out = net.translator.trunk["ipa_0"](node.cuda(), edge.cuda(), Rigid.from_tensor_7(rig.cuda()), nm.cuda())
    ref = O.ipa(params, "translator.trunk.ipa_0.", node, edge, rig[..., :4], rig[..., 4:], nm)
    valid = nm.bool()
    assert rel(out.cpu()[valid], ref[valid]) < 1e-4


@pytest.mark.parametrize("pair", PAIR_MODES)
def test_edge_transition_vs_oracle(golden_dir, params, pair):
    g = load(golden_dir, "net_forward_small.npz")
    f = small_feats(g)
    net = make_net(params, pair)
    nm = f["residue_mask"].float()
    node = g["node_embed"] * nm[..., None]
    edge = (g["edge_embed"] * (nm[..., None] * nm[..., None, :])[..., None]).bfloat16()
    eng = net.native("cuda")
    eng.reserve(2, 12, f["residue_idx"])
    out = eng.edge_transition(0, node.cuda().contiguous(), edge.cuda().contiguous(), nm.cuda().contiguous())
    ref = O.edge_transition(params, "translator.trunk.edge_transition_0.", node, edge.float())
    ref = ref * (nm[..., None] * nm[..., None, :])[..., None]
    assert rel(out.float(), ref) < 8e-3


@pytest.mark.parametrize("pair", PAIR_MODES)
def test_network_forward_vs_reference_golden(golden_dir, params, pair):
    g = load(golden_dir, "net_forward_small.npz")
    net = make_net(params, pair)
    with torch.no_grad():
        out = net(cuda(small_feats(g)), as_tensor_7=True)
    valid = small_feats(g)["residue_mask"].bool()
    assert rel(out["rigids"].cpu()[valid][:, 4:], g["out_rigids"][valid][:, 4:]) < 1e-4
    assert rel(out["rigids"].cpu()[valid][:, :4], g["out_rigids"][valid][:, :4]) < 1e-4
    assert rel(out["psi"].cpu()[valid], g["out_psi"][valid]) < 2e-3
    assert rel(out["atom37"].cpu()[valid][:, :5], g["out_atom37"][valid]) < 1e-4
    assert rel(out["atom14"].cpu()[valid][:, :5], g["out_atom14"][valid]) < 1e-4
    assert float(out["atom37"][..., 5:, :].abs().max()) == 0.0


def test_diffuser_vs_reference_golden(golden_dir):
    from str2str_b200.rigid import Rigid, Rotation

    g = load(golden_dir, "diffuser_steps.npz")
    d = make_diffuser()
    r0, rt, t, mask = g["r0"].cuda(), g["rt"].cuda(), g["t"], g["mask"].cuda()
    sc = d.score(Rigid.from_tensor_7(r0, normalize_quats=True), Rigid.from_tensor_7(rt), t, mask)
    assert sc["rot_score"].dtype == torch.float64
    assert rel(sc["rot_score"], g["rot_score"]) < 5e-5
    assert rel(sc["trans_score"], g["trans_score"]) < 1e-6
    diffuse = ((1 - g["fixed"]) * g["mask"]).cuda()
    ode = d.reverse(Rigid.from_tensor_7(rt), g["rot_score"].cuda(), g["trans_score"].cuda(), t, 0.02, diffuse_mask=diffuse).to_tensor_7()
    assert rel(ode, g["ode"]) < 2e-6
    sde = d.reverse(Rigid.from_tensor_7(rt), g["rot_score"].cuda(), g["trans_score"].cuda(), t, 0.02, diffuse_mask=diffuse,
                    noise_scale=0.7, probability_flow=False, rot_noise=g["z_rot"].cuda(), trans_noise=g["z_tr"].cuda()).to_tensor_7()
    assert rel(sde, g["sde"]) < 2e-6
    rig0 = Rigid(Rotation(rot_mats=O.quat_to_rotmat(g["r0"][..., :4]).cuda()), r0[..., 4:])
    fm = d.forward_marginal(rig0, t, diffuse_mask=diffuse, noise=(g["fm_axis"], g["fm_u"], g["fm_z"]))["rigids_t"]
    assert rel(fm, g["fm"]) < 2e-6


def test_sigma_buckets_bit_exact(golden_dir):
    g = load(golden_dir, "diffuser_steps.npz")
    d = make_diffuser()
    idx = d.rot_diffuser.t_to_idx(torch.linspace(0.01, 1.0, 397))
    assert torch.equal(idx, g["sigma_idx_grid"].long())
    assert np.array_equal(d.rot_diffuser.cdf_row(500), g["cdf_row_500"].numpy())


def _run_traj(golden_dir, params, name, pair, graph):
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    g = load(golden_dir, name)
    B, L, n, n_pad, n_fixed, seed = [int(v) for v in g["meta"]]
    feats = synthetic.make_features(1, L, seed=seed, n_pad=n_pad, n_fixed=n_fixed, random_aatype=True)
    net = make_net(params, pair)
    cfg = InferenceConfig(num_timesteps=2 * n, min_t=0.01)
    smp = ForwardBackwardSampler(net, make_diffuser(), cfg, use_cuda_graph=graph)
    q, x = synthetic.make_backbone(L, seed=seed)
    r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
    atom37, fin, psi = smp.forward_backward(cuda(feats), r0, 0.5, rigids_t=g["rigids_t"].cuda(), return_rigids=True)
    valid = synthetic.make_features(B, L, seed=seed, n_pad=n_pad, n_fixed=n_fixed)["residue_mask"].bool()
    ca, ca_ref = fin.cpu()[..., 4:][valid], g["final_rigids"][..., 4:][valid]
    r = rel(ca, ca_ref)
    print(f"{name} pair={pair} graph={graph}: C-alpha rel-L2 {r:.3e}, max|d| {float((ca - ca_ref).abs().max()):.3e} A, "
          f"launches {smp.launches}")
    assert r < 1e-4
    assert rel(torch.as_tensor(atom37)[valid][:, :5], g["final_atom37"][valid]) < 1e-4
    assert smp.launches > 0


@pytest.mark.parametrize("pair", PAIR_MODES)
@pytest.mark.parametrize("graph", [False, True])
def test_trajectory_cfg1_vs_reference_golden(golden_dir, params, pair, graph):
    """BASELINE.json configs[0]: 64 residues, 10 denoise steps, batch 1 — final C-alpha within 1e-4 relative."""
    _run_traj(golden_dir, params, "traj_cfg1_L64_n10.npz", pair, graph)


def test_trajectory_masked_vs_reference_golden(golden_dir, params):
    _run_traj(golden_dir, params, "traj_masked_L24_n6.npz", (1, 1), True)


@pytest.mark.parametrize("graph", [False, True])
def test_trajectory_L128_vs_reference_golden(golden_dir, params, graph):
    """128 residues x 20 denoise steps, 2 decoys with a padded tail: the tcgen05 pair kernels (L % 128 == 0), the panel GEMMs and
    the fused IPA kernel against a trajectory of the UNMODIFIED reference — final C-alpha within 1e-4 relative."""
    _run_traj(golden_dir, params, "traj_L128_n20.npz", (1, 1), graph)


@pytest.mark.parametrize("L", [128, 256, 384])
def test_pair_kernels_tc_vs_simt(params, L):
    """tcgen05 pair kernels against the SIMT restatement with identical rounding points, full forward."""
    B = 2
    feats = synthetic.make_features(B, L, seed=11, n_pad=5)
    q, x = synthetic.make_backbone(L, seed=11)
    g = torch.Generator().manual_seed(5)
    feats["rigids_t"] = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.2 * torch.randn(B, L, 7, generator=g)).float()
    feats["sc_ca_t"] = (x[None] + torch.randn(B, L, 3, generator=g)).float()
    feats["t"] = torch.tensor([0.4, 0.6])
    outs = []
    for pair in (0, 1):
        net = make_net(params, pair, 0)
        eng = net.native("cuda")
        fc = cuda(feats)
        eng.reserve(B, L, fc["residue_idx"])
        node, z = eng.embed(fc["t"], fc["residue_idx"], fc["fixed_mask"].float(), fc["sc_ca_t"], fc["residue_mask"].float())
        z2 = eng.edge_transition(0, node, z, fc["residue_mask"].float().contiguous())
        outs.append((z.float().cpu(), z2.float().cpu()))
    assert rel(outs[1][0], outs[0][0]) < 3e-3   # bf16 output rounding of slightly different fp32 sums
    assert rel(outs[1][1], outs[0][1]) < 3e-3, rel(outs[1][1], outs[0][1])


def test_se3_equivariance_full_size(params):
    """Size-independent property at BASELINE cfg-2 chain length: rotating + translating the input frames
    rotates + translates the predicted frames (IPA is SE(3)-equivariant), checked on the CUDA path at L=256."""
    B, L = 2, 256
    feats = synthetic.make_features(B, L, seed=13)
    q, x = synthetic.make_backbone(L, seed=13)
    feats["rigids_t"] = torch.cat([q, x], -1)[None].repeat(B, 1, 1).float()
    feats["sc_ca_t"] = torch.zeros(B, L, 3)
    feats["t"] = torch.tensor([0.3, 0.3])
    g = torch.Generator().manual_seed(3)
    qg = torch.nn.functional.normalize(torch.randn(4, generator=g), dim=0)
    Rg = O.quat_to_rotmat(qg)
    tg = torch.tensor([3.0, -2.0, 5.0])
    moved = dict(feats)
    rt = feats["rigids_t"].clone()
    rt[..., :4] = O.quat_mul(qg.expand(B, L, 4), rt[..., :4])
    rt[..., 4:] = rt[..., 4:] @ Rg.T + tg
    moved["rigids_t"] = rt
    net = make_net(params, 1)
    with torch.no_grad():
        a = net(cuda(feats), as_tensor_7=True)["rigids"].cpu()
        b = net(cuda(moved), as_tensor_7=True)["rigids"].cpu()
    expect = a[..., 4:] @ Rg.T + tg
    assert rel(b[..., 4:], expect) < 2e-5


@pytest.mark.parametrize("L", [64, 128])
def test_ipa_block_tensor_core_vs_oracle(params, L):
    """IPA block with the tensor-core projections / q.k^T / P.v (node_gemm=1) against the fp32 oracle."""
    from str2str_b200.rigid import Rigid

    B = 2
    g = torch.Generator().manual_seed(17)
    node = torch.randn(B, L, 256, generator=g)
    edge = torch.randn(B, L, L, 128, generator=g).bfloat16().float()
    q, x = synthetic.make_backbone(L, seed=17)
    rig = torch.cat([q, 0.1 * x], -1)[None].repeat(B, 1, 1) + 0.05 * torch.randn(B, L, 7, generator=g)
    nm = torch.ones(B, L)
    nm[:, -3:] = 0
    net = make_net(params, 1, 1)
    out = net.translator.trunk["ipa_1"](node.cuda(), edge.cuda(), Rigid.from_tensor_7(rig.cuda()), nm.cuda())
    ref = O.ipa(params, "translator.trunk.ipa_1.", node, edge, rig[..., :4], rig[..., 4:], nm)
    valid = nm.bool()
    r = rel(out.cpu()[valid], ref[valid])
    print(f"ipa tensor-core path L={L}: rel {r:.2e}")
    assert r < 5e-3  # single-pass bf16 projections / logits / P.v by design (tools/precision_probe.py)


def test_forward_tensor_core_vs_exact_L128(params):
    """Whole network forward: tcgen05 everywhere against SIMT pair kernels + exact fp32 node GEMMs."""
    B, L = 2, 128
    feats = synthetic.make_features(B, L, seed=19, n_pad=4, n_fixed=2)
    q, x = synthetic.make_backbone(L, seed=19)
    g = torch.Generator().manual_seed(2)
    feats["rigids_t"] = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.2 * torch.randn(B, L, 7, generator=g)).float()
    feats["sc_ca_t"] = (x[None] + torch.randn(B, L, 3, generator=g)).float()
    feats["t"] = torch.tensor([0.35, 0.35])
    outs = []
    for mode in ((0, 0), (1, 1)):
        net = make_net(params, mode)
        with torch.no_grad():
            outs.append(net(cuda(feats), as_tensor_7=True)["rigids"].cpu())
    valid = feats["residue_mask"].bool()
    r = rel(outs[1][valid][:, 4:], outs[0][valid][:, 4:])
    print(f"forward L=128 tensor-core vs exact: C-alpha rel {r:.2e}")
    assert r < 2e-5


def test_long_chain_forward_L512(params):
    """BASELINE cfg 4 chain length (512 residues): the whole forward runs (IPA slab = 128 KB of shared memory, 1 CTA/SM)
    and agrees between the tcgen05 path and the SIMT/exact path."""
    B, L = 1, 512
    feats = synthetic.make_features(B, L, seed=23)
    q, x = synthetic.make_backbone(L, seed=23)
    feats["rigids_t"] = torch.cat([q, x], -1)[None].repeat(B, 1, 1).float()
    feats["sc_ca_t"] = x[None].repeat(B, 1, 1).float()
    feats["t"] = torch.tensor([0.5])
    outs = []
    for mode in ((0, 0), (1, 1)):
        net = make_net(params, mode)
        with torch.no_grad():
            outs.append(net(cuda(feats), as_tensor_7=True)["rigids"].cpu())
    r = rel(outs[1][..., 4:], outs[0][..., 4:])
    print(f"forward L=512 tensor-core vs exact: C-alpha rel {r:.2e}")
    assert r < 2e-5


@pytest.mark.parametrize("L", [64, 128, 208, 256, 320, 384, 512])
def test_ipa_second_generation_vs_first_and_oracle(params, L):
    """Second-generation IPA path (point term folded into the logits GEMM, persistent tcgen05 pair kernel, split-bf16
    attention weights) against the first-generation kernels on the same engine and against the fp32 oracle.
    L=208 exercises a ragged second key block, L=64 a ragged first one; L > 256 takes the long-chain kernel (ring of
    128-key blocks): 320 = three blocks with a ragged last one, 384 = three full, 512 = four (BASELINE cfg 4 / cfg 5)."""
    B = 3 if L <= 256 else 2
    g = torch.Generator().manual_seed(29)
    node = torch.randn(B, L, 256, generator=g)
    edge = torch.randn(B, L, L, 128, generator=g).bfloat16()
    q, x = synthetic.make_backbone(L, seed=29)
    rig = torch.cat([q, 0.1 * x], -1)[None].repeat(B, 1, 1) + 0.05 * torch.randn(B, L, 7, generator=g)
    nm = torch.ones(B, L)
    nm[1, -5:] = 0
    nm[B - 1, :3] = 0
    net = make_net(params, 1, 1)
    eng = net.native("cuda")
    eng.reserve(B, L, torch.arange(L)[None].repeat(B, 1))
    outs = []
    for gen in (0, 1):
        eng.set_option("ipa_kernels", gen)
        outs.append(eng.ipa(2, node.cuda(), edge.cuda(), rig[..., :4].contiguous().cuda(), rig[..., 4:].contiguous().cuda(), nm.cuda()).cpu())
    ref = O.ipa(params, "translator.trunk.ipa_2.", node, edge.float(), rig[..., :4], rig[..., 4:], nm)
    valid = nm.bool()
    r01 = rel(outs[1][valid], outs[0][valid])
    r0, r1 = rel(outs[0][valid], ref[valid]), rel(outs[1][valid], ref[valid])
    print(f"ipa L={L}: gen2 vs gen1 {r01:.2e}; vs oracle gen1 {r0:.2e} gen2 {r1:.2e}")
    assert r01 < 2e-3 and r1 < 5e-3 and r1 < 2 * r0 + 1e-4


def test_sampler_graph_cache_across_shapes(params):
    """The sampler caches its static buffers and the captured iteration per (B, L): alternating chain lengths on one
    sampler (BASELINE cfg 5 style, un-padded batches of different L) must re-capture when the engine workspace is
    re-planned and must reproduce the first result bit for bit when the first shape comes back."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    net = make_net(params, 1, 1)
    smp = ForwardBackwardSampler(net, make_diffuser(), InferenceConfig(num_timesteps=8, min_t=0.01), use_cuda_graph=True)

    def run(L, B, seed):
        feats = synthetic.make_features(1, L, seed=seed)
        q, x = synthetic.make_backbone(L, seed=seed)
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(B, 1, 1).cuda(), normalize_quats=True)
        g = torch.Generator().manual_seed(seed)
        rt = (torch.cat([q, x], -1)[None].repeat(B, 1, 1) + 0.3 * torch.randn(B, L, 7, generator=g)).float()
        _, fin, _ = smp.forward_backward(cuda(feats), r0, 0.5, rigids_t=rt.cuda(), return_rigids=True)
        return fin.cpu()

    a1 = run(64, 2, 1)
    b1 = run(128, 3, 2)
    a2 = run(64, 2, 1)      # same shape as the first call: cached or re-captured, the result must not change
    b2 = run(128, 3, 2)
    a3 = run(64, 2, 1)
    assert torch.equal(a1, a2) and torch.equal(a1, a3) and torch.equal(b1, b2)
    assert torch.isfinite(a1).all() and torch.isfinite(b1).all()


def test_predict_step_delta_sweep_layout(params, tmp_path):
    """The caller of the path (predict_step, diffusion_module.py:214-369): directory layout, file names, model counts and
    the merged all_delta file for a 2-delta sweep with an odd last batch (3 replicas in batches of 2)."""
    from str2str_b200.predict import predict_step
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig

    L = 24
    net = make_net(params, 1, 1)
    cfg = InferenceConfig(delta_min=0.25, delta_max=0.5, delta_step=0.25, n_replica=3, replica_per_batch=2, num_timesteps=16, min_t=0.01)
    smp = ForwardBackwardSampler(net, make_diffuser(), cfg)
    feats = cuda(synthetic.make_features(1, L, seed=4, random_aatype=True))
    q, x = synthetic.make_backbone(L, seed=4)
    a, b, c, d = q.unbind(-1)
    R = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c), 2 * (b * c + a * d), a * a - b * b + c * c - d * d,
                     2 * (c * d - a * b), 2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1).reshape(L, 3, 3)
    M = torch.zeros(L, 4, 4)
    M[:, :3, :3], M[:, :3, 3], M[:, 3, 3] = R, x, 1
    feats["rigidgroups_gt_frames"] = M[None, :, None].repeat(1, 1, 8, 1, 1).cuda()
    feats["chain_index"] = torch.zeros(1, L, dtype=torch.long).cuda()
    feats["residue_index"] = (torch.arange(L)[None] + 1).cuda()
    feats["accession_code"] = ["toy"]
    out = predict_step(smp, feats, output_dir=str(tmp_path))
    assert out == os.path.join(str(tmp_path), "all_delta")
    for dname in ("0.25", "0.5"):
        txt = open(os.path.join(str(tmp_path), dname, "toy.pdb")).read()
        assert txt.count("MODEL") == 3 and txt.count("ENDMDL") == 3 and txt.endswith("END")
        assert all(len(ln) == 80 for ln in txt.split("\n")[:-1])
    merged = open(os.path.join(out, "toy.pdb")).read()
    assert merged.count("MODEL") == 6 and merged.count("ENDMDL") == 6 and merged.endswith("END".ljust(80) + "\n")


GEMM_CASES = [
    # M,   N,    K,    passes, relu, res,   outs          kernel / epilogue variant exercised
    (640, 320, 320, 3, 0, True, "c"),       # panel, K = 320 geometry (128 + 64 accumulators), residual
    (384, 672, 256, 3, 0, False, "c"),      # panel, ragged last chunk (32 columns)
    (200, 64, 128, 3, 1, False, "chl"),     # panel, M not a multiple of 128, ReLU, fp32 + both bf16 images
    (256, 768, 256, 1, 0, False, "h"),      # panel, single bf16 pass, bf16-only output (wide epilogue)
    (256, 960, 320, 3, 0, False, "hl"),     # panel, hi + lo images only (wide epilogue, alternating 128 / 64 chunks)
    (256, 256, 2688, 3, 0, True, "chl"),    # tile kernel (panel does not fit), residual + images
    (130, 36, 48, 3, 0, False, "c"),        # tile kernel, K % 64 != 0, N = 36
    (300, 256, 256, 3, 1, True, "c_inplace"),  # residual aliasing the output
    (148 * 128 + 300, 320, 320, 3, 0, True, "chl"),  # more panels than SMs: CTAs that own two panels (A restaged after a_free)
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=[f"M{c[0]}_N{c[1]}_K{c[2]}_p{c[3]}_{c[6]}" for c in GEMM_CASES])
def test_linear_tc_vs_fp64(case):
    """The tensor-core linear layer (panel kernel / tile kernel, every epilogue variant) through the C ABI against fp64:
    3-pass split-bf16 within 3e-5 relative L2 (the dropped lo*lo term is ~2^-16), single bf16 pass within 1e-2; the bf16 images
    are the rounded result (hi) and its remainder (hi + lo within 2^-15 of the fp32 output)."""
    from str2str_b200 import _lib

    M, N, K, passes, relu, use_res, outs = case
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g) if use_res else None
    ref = A.double() @ W.double().T + b.double()
    if relu:
        ref = torch.relu(ref)
    if use_res:
        ref = ref + R.double()
    Ad, Wd, bd = A.cuda(), W.cuda(), b.cuda()
    inplace = outs == "c_inplace"
    kinds = outs.split("_")[0]
    Cd = R.cuda().clone() if inplace else (torch.full((M, N), float("nan"), device="cuda") if "c" in kinds else None)
    Rd = Cd if inplace else (R.cuda() if use_res else None)
    Hd = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if "h" in kinds else None
    Ld = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16) if "l" in kinds else None
    p = lambda t: _lib.ptr(t, None) if t is not None else None  # fp32 inputs / result, bf16 images
    _lib.check(lib.s2s_linear_tc(p(Ad), p(Wd), p(bd), p(Rd), p(Cd), p(Hd), p(Ld), M, N, K, passes, relu, _lib.stream()))
    torch.cuda.synchronize()
    tol = 3e-5 if passes == 3 else 1e-2
    if Cd is not None:
        assert torch.isfinite(Cd).all()
        assert rel(Cd, ref) < tol, rel(Cd, ref)
    if Hd is not None:
        assert rel(Hd.float(), ref) < max(tol, 4e-3)            # bf16 rounding of the output
        if Cd is not None:
            assert torch.equal(Hd, Cd.bfloat16())
    if Ld is not None:
        assert rel(Hd.float() + Ld.float(), ref) < max(tol, 4e-5)


@pytest.mark.parametrize("graph", [False, True])
def test_trajectory_scheduler_equals_per_delta_sampler(params, graph):
    """Continuous batching (str2str_b200/scheduler.py): trajectories of two deltas (6 and 10 denoising steps) stream through a
    2-row persistent batch; every trajectory must come out as from ForwardBackwardSampler.forward_backward run per delta on the
    same perturbed start frames (decoys do not interact on the path, so the values agree to fp32 rounding noise)."""
    from str2str_b200.rigid import Rigid
    from str2str_b200.sampler import ForwardBackwardSampler, InferenceConfig
    from str2str_b200.scheduler import TrajectoryScheduler

    L = 128
    feats = synthetic.make_features(1, L, seed=21, n_pad=3, random_aatype=True)
    q, x = synthetic.make_backbone(L, seed=21)
    gt = torch.zeros(1, L, 8, 4, 4)
    gt[0, :, 0, :3, :3] = O.quat_to_rotmat(torch.nn.functional.normalize(q, dim=-1))
    gt[0, :, 0, :3, 3] = x
    gt[0, :, 0, 3, 3] = 1.0
    batch = cuda(dict(feats, rigidgroups_gt_frames=gt))
    net = make_net(params, 1)
    cfg = InferenceConfig(num_timesteps=20, min_t=0.01)
    smp = ForwardBackwardSampler(net, make_diffuser(), cfg, use_cuda_graph=graph)
    work = [(0.3, 3), (0.5, 2)]
    gen = torch.Generator().manual_seed(9)
    start = {}
    for d, n in work:
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(n, 1, 1).cuda(), normalize_quats=True)
        noise = (torch.randn(n, L, 3, generator=gen), torch.rand(n, L, generator=gen), torch.randn(n, L, 3, generator=gen))
        start[d] = smp.diffuser.forward_marginal(r0, d * torch.ones(n), diffuse_mask=torch.ones(n, L, dtype=torch.float64), noise=noise)["rigids_t"]
    sched = TrajectoryScheduler(smp, slots=2)
    atom37, rig = sched.run(batch, work, rigids_t=start, return_rigids=True)
    assert sched.iterations == 25 and sched.row_iterations == 3 * 7 + 2 * 11   # row 0: 7 + 7 + 11 phases, row 1: 7 + 11 (1 + n each)
    for d, n in work:
        r0 = Rigid.from_tensor_7(torch.cat([q, x], -1)[None].repeat(n, 1, 1).cuda(), normalize_quats=True)
        a_ref, rig_ref, _ = smp.forward_backward(batch, r0, d, rigids_t=start[d], return_rigids=True)
        # measured on B200: 1.7e-6 (fp32 reordering noise of a different batch composition; the parity gate is 1e-4)
        assert rel(rig[d][..., 4:], rig_ref[..., 4:]) < 1e-5, (d, rel(rig[d][..., 4:], rig_ref[..., 4:]))
        assert np.abs(atom37[d] - a_ref).max() < 1e-3
