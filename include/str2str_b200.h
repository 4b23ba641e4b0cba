/*
 * str2str_b200 — C ABI of the B200-native Str2Str denoising hot path.
 *
 * The reference (lujiarui/Str2Str) is pure Python: it has no FFI.  Its "operator interface" for this path is a
 * set of Python call signatures.  This header is the boundary those signatures are re-implemented
 * on: plain device pointers, sizes and a cudaStream_t (passed as void*), no torch types.  The Python host mirror
 * (str2str_b200/net, str2str_b200/score) binds it with ctypes; INTEGRATION.md shows the stub a reference
 * maintainer would add.  Every entry point returns 0 on success and a non-zero code on failure, in which case
 * s2s_last_error() describes the problem; nothing aborts and nothing falls back to the CPU.
 *
 * All tensors are contiguous, row-major, on the current device.  fp32 unless noted; residue indices and
 * residue types are int64 as in the reference batch dict.  B = decoys in flight, L = residues.
 */
#ifndef STR2STR_B200_H
#define STR2STR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s2s_ctx s2s_ctx;

#define S2S_ABI_VERSION 1

int s2s_abi_version(void);
const char* s2s_last_error(void);

/* ---- context: weights + workspace ------------------------------------------------------------------------
 * Replaces DenoisingNet.__init__ + load_state_dict (reference src/models/net/denoising_ipa.py:162-169,
 * src/utils/checkpoint_utils.py:16-20).  Host tables (all host pointers, copied):
 *   tfreq[16]      exp(-k ln(1e4)/15)                 denoising_ipa.py:39-41
 *   pdenom[16]     2056^(2k/32)                        denoising_ipa.py:26-30
 *   bin_lower[22]  linspace(1e-5, 20, 22)              geo_utils.py:49-53
 *   backbone[21*33] ideal backbone geometry per residue type (str2str_b200/backbone_constants.py)
 */
s2s_ctx* s2s_create(const float* tfreq, const float* pdenom, const float* bin_lower, const float* backbone);
void s2s_destroy(s2s_ctx* ctx);
/* Register one tensor of the reference state dict by its key (e.g. "translator.trunk.ipa_0.linear_q.weight").
 * `data` is a DEVICE pointer to fp32 that must stay alive until s2s_destroy. */
int s2s_set_param(s2s_ctx* ctx, const char* name, const float* data, int64_t numel);
/* Check that all 274 tensors are present with the right sizes and build the derived device tensors
 * (bf16 / split / transposed weight images, concatenated projections, softplus head weights). */
int s2s_finalize(s2s_ctx* ctx, void* stream);
/* Options: "pair_kernels" 1 = tcgen05 pair kernels (default), 0 = SIMT cross-check kernels (tests);
 *          "node_gemm"    1 = tensor-core GEMMs where the parity budget allows (default), 0 = exact fp32 FFMA everywhere (tests);
 *          "ipa_kernels"  1 = second-generation IPA path (default; needs node_gemm = 1 and L <= 512, longer chains fall back
 *                         automatically): point-attention term folded into the logits GEMM, persistent TMA + tcgen05 pair kernel,
 *                         split-bf16 attention weights; 0 = first-generation kernels (A/B, tests);
 *          "et_pair"      1 = EdgeTransition on CTA pairs (tcgen05 cta_group::2: M = 256 MMAs issued once per pair of SMs, half of the
 *                         weight stream per SM; default; measured 2.22 vs 2.49 ms per launch at cfg2, same values bit for bit),
 *                         0 = one CTA per row tile (A/B, tests);
 *          "embed_table"  1 = the edge embedder builds each decoy's (fixed_i, fixed_j, index offset, distogram bin) -> embedding
 *                         table and expands it into the pair tensor whenever that table is well below L^2 rows (default; same
 *                         values bit for bit), 0 = always run the MLP on every pair row (A/B, tests);
 *          "chain"        the row-local layers of the node track between two attention kernels (sequence-transformer out_proj /
 *                         norm1 / linear1 / linear2 / norm2, the next in_proj or the post-transformer linear, NodeTransition,
 *                         the per-residue terms of the EdgeTransition, the torsion head) as ONE launch each (gemm_chain.cu,
 *                         LayerNorm as a step epilogue; 162 -> 88 launches per forward, same values bit for bit):
 *                         1 = when the 128-row panels outnumber half the SMs (default: below that separate launches, which then split their output
 *                         columns over the idle SMs, are faster), 2 = always, 0 = never. */
int s2s_set_option(s2s_ctx* ctx, const char* key, int value);
/* Size the workspace for (B, L) and build the relative-position table for residue-index offsets in
 * [d_min, d_max] (= min/max of residue_idx[i] - residue_idx[j]); d_max < d_min keeps the table already planned
 * (shape-only call).  Must be called before the forward entry points and outside CUDA-graph capture; the
 * allocation only ever grows (a larger B, L or offset range re-allocates; smaller shapes reuse it).
 * ANY chain length L >= 1 is accepted by every entry point: the tensor-core kernels tile chains in units of 32
 * residues, other lengths are padded inside the library with residues that take part in no reduction (mask 0
 * for IPA / pair rows, removed from the sequence transformer's keys), so results equal the un-padded computation
 * of the reference; offsets outside the planned range are clamped to the table (never read out of bounds). */
int s2s_reserve(s2s_ctx* ctx, int B, int L, int d_min, int d_max, void* stream);

/* ---- score network ------------------------------------------------------------------------------------------
 * DenoisingNet.forward (denoising_ipa.py:171-211) without the backbone build:
 *   rigids_t [B,L,7] (quat wxyz + trans in Angstrom), sc_ca [B,L,3], t [B], residue_idx [B,L] int64,
 *   residue_mask / fixed_mask [B,L], gt_psi [B,L,2] (torsion_angles_sin_cos[..., 2, :])
 *   -> out_rigids [B,L,7], out_psi [B,L,2]. */
int s2s_net_forward(s2s_ctx* ctx, int B, int L, const float* rigids_t, const float* sc_ca, const float* t,
                    const int64_t* residue_idx, const float* residue_mask, const float* fixed_mask,
                    const float* gt_psi, float* out_rigids, float* out_psi, void* stream);

/* Stage entry points (the same code s2s_net_forward runs), exposed for module-level parity tests.
 * EmbeddingModule.forward (denoising_ipa.py:107-159) followed by the mask multiply (:186-187):
 *   node_out [B,L,256] fp32, z_out [B,L,L,128] bf16. */
int s2s_embed(s2s_ctx* ctx, int B, int L, const float* t, const int64_t* residue_idx, const float* fixed_mask,
              const float* sc_ca, const float* residue_mask, float* node_out, void* z_out, void* stream);
/* TranslationIPA.forward (ipa.py:331-387) on given embeddings: node_embed [B,L,256] fp32 and z [B,L,L,128] bf16,
 * both already multiplied by their masks.  gt_psi may be NULL: out_psi is then the raw torsion head output. */
int s2s_trunk(s2s_ctx* ctx, int B, int L, const float* node_embed, const void* z, const float* rigids_t,
              const float* residue_mask, const float* fixed_mask, const float* gt_psi, float* out_rigids,
              float* out_psi, void* stream);
/* InvariantPointAttention.forward of trunk block `blk` (ipa.py:100-268): quat [B,L,4] (not normalised),
 * trans_nm [B,L,3] (already x0.1), z [B,L,L,128] bf16 -> out [B,L,256] (before the mask multiply of ipa.py:350). */
int s2s_ipa(s2s_ctx* ctx, int blk, int B, int L, const float* node, const void* z, const float* quat,
            const float* trans_nm, const float* residue_mask, float* out, void* stream);
/* EdgeTransition.forward of block `blk` followed by the edge-mask multiply (layers.py:170-185, ipa.py:371-372);
 * z_out may alias z_in. */
int s2s_edge_transition(s2s_ctx* ctx, int blk, int B, int L, const float* node, const void* z_in,
                        const float* residue_mask, void* z_out, void* stream);

/* Module-level pieces of the trunk on `rows` residue rows [rows,256] (rows <= reserved B * L), for drop-in use of the
 * reference's sub-modules: NodeTransition.forward (layers.py:138-145) -> out [rows,256]; TorsionAngleHead.forward
 * (layers.py:199-213) -> out [rows,2]; BackboneUpdate.forward (layers.py:232-241) -> out [rows,6]. */
int s2s_node_transition(s2s_ctx* ctx, int blk, int64_t rows, const float* s, float* out, void* stream);
int s2s_torsion_head(s2s_ctx* ctx, int64_t rows, const float* s, float* out, void* stream);
int s2s_backbone_update(s2s_ctx* ctx, int blk, int64_t rows, const float* s, float* out, void* stream);

/* ---- SE(3) diffusion step ------------------------------------------------------------------------------------
 * FrameDiffuser.score + FrameDiffuser.reverse (reference src/models/score/frame.py:109-210).
 * Per-decoy schedule scalars are computed by the host exactly as the reference computes them (same torch ops,
 * so the integer sigma bucket is bit-identical) and passed in:
 *   sched_f [B][8] = t, sigma_q = discrete_sigma[bucket], g_rot(t), g_rot(t)^2, exp(-beta(t)/2), 1-exp(-beta(t)),
 *                    b(t), sqrt(b(t))          (so3.py:205-234, r3.py:26-41)
 *   sched_d [B][2] = dt, sqrt(dt)              (double)
 * mode 0: fused score + reverse (scores optional outputs); 1: scores only; 2: reverse from given scores.
 * Scores are fp64 like the reference's (its fp64 masks promote them, frame.py:137-138).
 * rot_noise / trans_noise [B,L,3]: N(0,1) draws, required when probability_flow == 0. */
int s2s_se3_step(int B, int L, const float* rigids_t, const float* rigids_0, const float* residue_mask,
                 const float* diffuse_mask, const float* sched_f, const double* sched_d, const float* rot_noise,
                 const float* trans_noise, float noise_scale, int probability_flow, int mode, double* rot_score,
                 double* trans_score, float* rigids_out, void* stream);
/* FrameDiffuser.forward_marginal (frame.py:36-107) with the three random draws supplied by the caller:
 *   rot0 [B,L,3,3], trans0 [B,L,3], sched_f [B][2] = exp(-beta/2), sqrt(1-exp(-beta)), cdf [B][1000] fp64 row of
 *   SO3Diffuser._cdf for each decoy's sigma bucket, omega_grid [1000]. */
int s2s_se3_perturb(int B, int L, const float* rot0, const float* trans0, const float* diffuse_mask,
                    const float* sched_f, const double* cdf, const float* omega_grid, const float* axis_noise,
                    const float* u_noise, const float* trans_noise, float* rigids_out, void* stream);
/* Random draws keyed by GLOBAL decoy id (replaces the torch.randn / torch.rand calls of so3.py:259-262, r3.py:66,109):
 * out [B][n_per_decoy] fp32; decoy b of the call is decoy first_decoy + b of the job and reads Philox4x32-10 subsequence
 * (first_decoy + b) of `seed`, block `stream_id` (0 axis, 1 angle quantile, 2 translation of the perturbation; 16 + 2k / 17 + 2k
 * rotation / translation noise of SDE iteration k).  uniform = 0: N(0,1); 1: U[0,1).  A decoy's draws therefore do not depend
 * on the batch, rank or world size it is sampled in. */
int s2s_rng_fill(float* out, int B, int64_t n_per_decoy, uint64_t seed, int64_t first_decoy, int stream_id, int uniform, void* stream);
/* The same draws for rows that belong to different decoys / iterations (continuous batching of trajectories): row b draws for
 * decoy decoy_ids[b] (device int64 [B]; a negative id leaves the row untouched), block stream_ids[b] (device int32 [B]). */
int s2s_rng_fill_rows(float* out, int B, int64_t n_per_decoy, uint64_t seed, const int64_t* decoy_ids, const int32_t* stream_ids,
                      int uniform, void* stream);
/* compute_backbone (reference src/common/all_atom.py:141-173): atom37 [B,L,37,3], atom14 [B,L,14,3] (nullable).
 * aatype may be NULL (all alanine, as the reference does for aatype=None). */
int s2s_backbone_atoms(s2s_ctx* ctx, int rows, const float* rigids, const float* psi, const int64_t* aatype,
                       float* atom37, float* atom14, void* stream);

/* Exact-fp32 GEMM  C[M,N] = A[M,K] W[N,K]^T + bias  (unit test hook for the node-track GEMM). */
int s2s_linear_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int relu,
                   void* stream);

/* Tensor-core linear layer of the node track (reference nn.Linear call sites: src/models/net/ipa.py:131-158,262,306-318;
 * layers.py:138-145,199-213):  Y = act(A W^T + bias) [+ res]  with A [M,K], W [N,K] (row-major fp32, device).  passes = 3: split-bf16
 * (hi/lo) operands, three tcgen05 MMA passes, fp32 accumulate (~16 mantissa bits per operand); passes = 1: single bf16.  Outputs
 * (any subset, at least one): C [M,N] fp32, out_hi / out_lo [M,N] bf16 images of the result (out_lo needs out_hi).  Routed to the
 * panel kernel (A resident in tensor memory) when K % 64 == 0 and the panel fits, else to the tile kernel.  K % 16 == 0, N % 4 == 0.
 * `res` may alias `C`.  Temporaries (operand images) are allocated and freed on `stream`. */
int s2s_linear_tc(const float* A, const float* W, const float* bias, const float* res, float* C, void* out_hi, void* out_lo,
                  int M, int N, int K, int passes, int relu, void* stream);

/* Per-kernel device timing for the benchmark's roofline line: when enabled, CUDA events bracket the launches of
 * the named kernels ("edge_transition", "edge_embed", "ipa_pair_attention", "gemm") on their stream.  Must be
 * off during CUDA-graph capture.  s2s_profile_read returns 1 if no launch of `name` was recorded. */
void s2s_profile_enable(int on);
void s2s_profile_reset(void);
int s2s_profile_read(const char* name, double* total_ms, int64_t* count);
/* all recorded names as "name\ttotal_ms\tcount\n" lines; returns the length, or -(needed capacity) if buf is too small */
int s2s_profile_list(char* buf, int cap);

/* number of kernels launched by this library since load (for bench.py's gpu_launches) */
int64_t s2s_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* STR2STR_B200_H */
