// Internal launcher declarations shared by the .cu files (not part of the C ABI; see include/str2str_b200.h).
#pragma once
#include "common.cuh"

namespace s2s {

// ---- gemm.cu ------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A = nullptr; long lda = 0; long sAb = 0, sAh = 0;
  const float* B = nullptr; long ldb = 0; long sBb = 0, sBh = 0;
  float* C = nullptr;       long ldc = 0; long sCb = 0, sCh = 0;
  const float* bias = nullptr;
  const float* res = nullptr; long ldres = 0;  // residual shares C's batch offsets
  const float* row_pre = nullptr;               // [M] scale on the accumulator (before bias)
  const float* row_post = nullptr;              // [M] scale after bias/activation
  int M = 0, N = 0, K = 0;
  int nb = 1, nh = 1;
  int b_kn = 0;   // 0: B is [N][K] (nn.Linear weight); 1: B is [K][N]
  int relu = 0;
  float alpha = 1.f;
};
void gemm_f32(const GemmArgs& g, cudaStream_t st);
// tensor-core GEMM (gemm_tc.cu): C = epilogue(A B^T), bf16 K-major operands via TMA, fp32 accumulate in TMEM.
// Operand tensors are [rows][cols] bf16 with a row pitch; the box of tile (m0 | n0), batch (ib, ih), k-block kb is at
// row = ib*rb + ih*rh + m0, col = ib*cb + ih*ch + kb*64.  passes = 1 (plain bf16) or 3 (hi/lo split, needs *_lo).
struct TcGemm {
  const bf16 *A_hi = nullptr, *A_lo = nullptr; size_t a_rows = 0, a_cols = 0, a_pitch = 0; int a_cb = 0, a_ch = 0, a_rb = 0, a_rh = 0;
  const bf16 *B_hi = nullptr, *B_lo = nullptr; size_t b_rows = 0, b_cols = 0, b_pitch = 0; int b_cb = 0, b_ch = 0, b_rb = 0, b_rh = 0;
  int M = 0, N = 0, K = 0, nb = 1, nh = 1, passes = 1, relu = 0, vt_L = 0;
  float alpha = 1.f;
  const float *bias = nullptr, *row_pre = nullptr, *row_post = nullptr, *res = nullptr; long ldres = 0;
  float* C = nullptr; long ldc = 0, sCb = 0, sCh = 0;
  bf16 *out_hi = nullptr, *out_lo = nullptr; long ldo = 0;  // optional bf16 (split) copies of the result; batched calls
                                                            // index them like C (sCb / sCh), so ldo must equal ldc there
  // optional transposed (per head) bf16 copy of the 'v' columns of a fused q|k|v projection: column n is v of head
  // h = (n - vt_col0) / vt_stride, channel c = (n - vt_col0) % vt_stride - vt_off when 0 <= c < vt_width
  bf16 *out_vt = nullptr, *out_vt_lo = nullptr;
  int vt_col0 = 2048, vt_stride = 512, vt_off = 256, vt_width = 256, vt_heads = 8;
  // optional second K segment (passes == 1): C += A2 B2^T over K2 more columns, box coordinates like A / B
  const bf16 *A2 = nullptr, *B2 = nullptr; int K2 = 0;
  size_t a2_rows = 0, a2_cols = 0, a2_pitch = 0; int a2_cb = 0, a2_ch = 0, a2_rb = 0, a2_rh = 0;
  size_t b2_rows = 0, b2_cols = 0, b2_pitch = 0; int b2_cb = 0, b2_ch = 0, b2_rb = 0, b2_rh = 0;
  long bias_sb = 0, bias_sh = 0;  // batch strides of `bias` (per-(batch, head) column bias when non-zero)
  // b_mn = 1: B is given as [K rows][N columns] (N contiguous, e.g. the v columns of a row-major q|k|v buffer) and is read
  // as an MN-major UMMA operand: box row = ib*b_rb + ih*b_rh + kb*64, box column = ib*b_cb + ih*b_ch + n0 (+64).
  int b_mn = 0;
  int force_panel = 0;  // batched call that must take the panel kernel (gemm_tc_splitk)
};
void gemm_tc(const TcGemm& g, cudaStream_t st);
void gemm_tc_splitk(const TcGemm& g, float* scratch, cudaStream_t st);  // scratch: ceil(K / 320) * M * N floats
void split_bf16(const float* src, long ld, int rows, int cols, bf16* hi, bf16* lo, cudaStream_t st);

// ---- gemm_chain.cu: consecutive row-local linear layers (split-bf16, 3 passes) in one launch ----------------------------
// step:  y = act(A W^T + bias) * row_post + res  ->  C (fp32) and / or its split-bf16 image (out_hi, out_lo);
//        with ln_w set, C is a scratch and LayerNorm(C) * ln_w + ln_b, times ln_scale[row], goes to ln_out (+ ln_hi / ln_lo).
// A of every step is a split-bf16 image [M][K] (pitch lda, 0 = K): the caller's for step 0, an earlier step's output after.
constexpr int CHAIN_MAX_STEPS = 10;
struct ChainStep {
  const bf16 *A_hi = nullptr, *A_lo = nullptr; long lda = 0;
  const bf16 *W_hi = nullptr, *W_lo = nullptr; long ldw = 0;  // [N][K] (nn.Linear weight), row pitch ldw
  int N = 0, K = 0, relu = 0;
  const float *bias = nullptr, *res = nullptr, *row_post = nullptr; long ldres = 0;
  float* C = nullptr; long ldc = 0;
  bf16 *out_hi = nullptr, *out_lo = nullptr; long ldo = 0;
  const float *ln_w = nullptr, *ln_b = nullptr, *ln_scale = nullptr;
  float* ln_out = nullptr; bf16 *ln_hi = nullptr, *ln_lo = nullptr; long ld_ln = 0;  // one pitch for ln_out and its images
};
void gemm_chain(const ChainStep* steps, int n_steps, int M, cudaStream_t st);

// ---- rows.cu ------------------------------------------------------------------------------------------
// y has pitch D.  LnExtra: the image (y_hi, y_lo) and an optional second fp32 copy y2 share the pitch ld2 (0 = D); `tail`
// ([rows][tail_w], pitch tail_ld) is copied into columns [D, D + tail_w) of y2 and of the image.
struct LnExtra {
  int ld2 = 0;
  float* y2 = nullptr;
  const float* tail = nullptr;
  int tail_ld = 0, tail_w = 0;
};
void layernorm(const float* x, const float* res, const float* w, const float* b, const float* rowscale, float* y,
               int rows, int D, cudaStream_t st, bf16* y_hi = nullptr, bf16* y_lo = nullptr, const LnExtra& ex = LnExtra());
void softmax_keybias(float* S, const float* keybias, int nb, int nh, int L, cudaStream_t st, bf16* P_hi = nullptr,
                     bf16* P_lo = nullptr);
void node_features(const float* t, const long long* ridx, const float* fixed, const float* tfreq,
                   const float* pdenom, float* feat, float* tf, int B, int L, cudaStream_t st);
void relpos_features(const float* pdenom, float* out, int d_min, int n, cudaStream_t st);
void psi_finalize(const float* u, const float* gt_psi, const float* fixed, float* psi, int rows, cudaStream_t st);
void concat_skip(const float* node, const float* skip, float* out, long rows, cudaStream_t st, bf16* out_hi = nullptr,
                 bf16* out_lo = nullptr);
void make_masks(const float* rmask, const float* fixed, const float* hard, float* diffuse, float* keybias, int n, cudaStream_t st);
// internal chain-length padding (rows.cu): per-residue inputs [B][L] -> [B][Lp]; any source may be null (skipped)
struct PadInputs {
  int B, L, Lp;
  const float *rig = nullptr, *sc = nullptr, *rmask = nullptr, *fixed = nullptr, *psi = nullptr;
  const long long* ridx = nullptr;
  float *o_rig = nullptr, *o_sc = nullptr, *o_rmask = nullptr, *o_fixed = nullptr, *o_psi = nullptr, *o_hard = nullptr;
  long long* o_ridx = nullptr;
};
void pad_inputs(const PadInputs& a, cudaStream_t st);
void repitch_rows(const float* src, float* dst, int B, int Ls, int Ld, int W, cudaStream_t st);  // [B][Ls][W] -> [B][Ld][W]
void repitch_pair(const bf16* src, bf16* dst, int B, int Ls, int Ld, cudaStream_t st);           // [B][Ls][Ls][128] -> [B][Ld][Ld][128]

// ---- pair kernels -------------------------------------------------------------------------------------
// distogram bin of |a-b| with the reference's strict inequalities (geo_utils.py:44-56); -1 = no bin
__device__ __forceinline__ int pair_distogram_bin(const float* a, const float* b, const float* lower) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
#pragma unroll 1
  for (int k = N_BINS - 1; k >= 0; --k) {
    if (d > lower[k]) {
      const float upper = (k < N_BINS - 1) ? lower[k + 1] : 1e8f;
      return d < upper ? k : -1;
    }
  }
  return -1;
}

struct EdgeEmbedArgs {
  int B, L, d_min;
  int n_off = 1;           // rows of Tpos: offsets outside [d_min, d_min + n_off) are clamped (never read out of bounds)
  const float* Ti;    // [B*L][128]  W1[:, 0:33] tf_i + b1
  const float* Tj;    // [B*L][128]  W1[:,33:66] tf_j
  const float* Tpos;  // [n_off][128] W1[:,66:98] pos(d_min + r)
  const float* Wd;    // [22][128]   W1[:,98+bin]
  const float* bin_lower;  // [22]
  const float* sc_ca;      // [B*L][3] Angstrom
  const long long* ridx;   // [B*L]
  const float* mask;       // [B*L]
  const bf16 *W2, *W3;     // [128][128] (out, in)
  const bf16 *W2t, *W3t;   // [in][out]
  const float *b2, *b3, *ln_w, *ln_b;
  bf16* z_out;             // [B,L,L,128]
  const bf16* wimg = nullptr;  // tcgen05 path: pre-swizzled weight blocks (build_ee_wimg)
  const float* vec4 = nullptr; // [b2 | b3 | ln_w | ln_b] packed, device (pipelined kernel: copied into its constant bank)
  // embedding table (pair_tc4.cu, MODE 2): 0 = direct kernel only, 1 = table without fixed-residue variants, 4 = with
  int table_variants = 0;
  const float* fixed = nullptr;   // [B*L] fixed mask (the table's row classes)
  bf16* table = nullptr;          // [B*4][tab_rows_per_block(n_off)][128]
  int* tab_ctl = nullptr;         // edge_embed_ctl_ints(B) ints
  unsigned char* cls = nullptr;   // [B*L]
};
size_t edge_embed_table_elems(int B, int n_off, int variants);
size_t edge_embed_ctl_ints(int B);
void edge_embed_simt(const EdgeEmbedArgs& a, cudaStream_t st);
void edge_embed_tc2(const EdgeEmbedArgs& a, cudaStream_t st);  // pipelined tcgen05 kernel, pair_tc4.cu

struct EdgeTransitionArgs {
  int B, L;
  const bf16* z_in;   // [B,L,L,128]
  const float* u;     // [B*L][384] W1[:,128:256] n'_i + b1
  const float* v;     // [B*L][384] W1[:,256:384] n'_j
  const float* p;     // [B*L][128] Wf[:,128:256] n'_i + bf
  const float* q;     // [B*L][128] Wf[:,256:384] n'_j
  const float* mask;  // [B*L]
  const bf16 *W1z, *W2, *Wfh, *Wfz;      // [384][128], [384][384], [128][384], [128][128]  (out, in)
  const bf16 *W1zt, *W2t, *Wft, *Wfzt;   // transposed copies [in][out]
  const float *b2, *ln_w, *ln_b;
  bf16* z_out;        // may alias z_in
  const bf16* nprime_bf16 = nullptr;  // tcgen05 path: n' [B*L][128] in bf16 (the n'_j operand rows)
  const bf16* wimg3 = nullptr;        // tcgen05 path: pre-swizzled weight blocks in consumption order (build_et3_wimg)
  int wimg_copies = 1;                // number of identical images laid out back to back (L2 hot-spot relief)
};
void edge_transition_simt(const EdgeTransitionArgs& a, cudaStream_t st);
void edge_transition_tc3(const EdgeTransitionArgs& a, cudaStream_t st);
void edge_transition_pair(const EdgeTransitionArgs& a, cudaStream_t st);  // pair_tc5.cu: CTA pairs, cta_group::2 MMAs
bool edge_transition_pair_supported(int B, int L);
size_t et3_wimg_elems();
void build_et3_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st);
size_t ee_wimg_elems();
void build_ee_wimg(const float* W2, const float* W3, bf16* dst, cudaStream_t st);
void f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t st);

// ---- ipa.cu -------------------------------------------------------------------------------------------
constexpr int VP_PITCH = 40;  // value-point columns per head in the split-bf16 images: 36 + 4 zeros (16-byte aligned TMA boxes)
constexpr int PT_K = 80;  // point columns of the fused logits GEMM: 3 x 24 split-bf16 coordinates + 8 zeros
struct IpaPointsAug {     // optional tensor-core operand outputs of ipa_points (see the kernel)
  bf16 *qp_aug = nullptr, *kp_aug = nullptr;  // [rows][N_H][PT_K]
  float* colbias = nullptr;                   // [B][N_H][L]
  bf16 *vp_hi = nullptr, *vp_lo = nullptr;    // [rows][N_H][VP_PITCH]
  const float* pt_w = nullptr;                // [N_H] softplus(head_weights) * sqrt(1/108)
  float inv_alpha = 1.f;                      // 1 / (alpha of the logits GEMM): the point columns are pre-divided by it
  int L = 0;
};
void ipa_points(const float* qp_raw, long ld_q, const float* kvp_raw, long ld_kv, const float* quat,
                const float* trans, float* q_pts, float* k_pts, float* v_pts, int rows, cudaStream_t st,
                const IpaPointsAug& aug = IpaPointsAug());
void ipa_point_logits(float* S, const float* q_pts, const float* k_pts, const float* pt_w, int B, int L,
                      cudaStream_t st);
struct IpaPairArgs {
  int B, L;
  const bf16* z;       // [B,L,L,128]
  float* S;            // [B,H,L,L]  in: scaled q.k + point term; out: attention weights
  const float* mask;   // [B*L]
  const bf16 *Wb_hi, *Wb_lo;  // linear_b weight split in two bf16 terms, [8][128]
  const float* bb;     // [8]
  const float* Wdz_t;  // down_z weight transposed, [128][32]
  const float* bdz;    // [32]
  float* o_pair;       // [B*L] rows, H*32 wide, row stride ld_opair
  long ld_opair;
  bf16* P_bf16 = nullptr;  // optional bf16 copy of the attention weights [B,H,L,L] (A operand of the tensor-core P.V)
  // second-generation kernel (ipa_tc.cu): P leaves as split bf16 (P_bf16 = hi part, P_lo), S is read only
  bf16* P_lo = nullptr;
  const bf16* wb_img = nullptr;  // build_ipa_wb_img
  bf16 *opair_hi = nullptr, *opair_lo = nullptr;  // optional split-bf16 image of o_pair (same indexing); replaces the fp32 write
};
void ipa_pair_attention(const IpaPairArgs& a, cudaStream_t st);
void ipa_pair_attention_tc(const IpaPairArgs& a, cudaStream_t st);
bool ipa_pair_attention_tc_supported(int L);
size_t ipa_wb_img_elems();
void build_ipa_wb_img(const float* Wb, bf16* dst, cudaStream_t st);
void ipa_finalize_points(const float* opt_glob, const float* quat, const float* trans, float* feats, int rows,
                         cudaStream_t st, bf16* f_hi = nullptr, bf16* f_lo = nullptr);
void softplus_point_weights(const float* head_w, float* pt_w, cudaStream_t st);

// ---- tfm_attn.cu --------------------------------------------------------------------------------------
bool tfm_attention_supported(int L);
void tfm_attention(const bf16* qkv, const float* keybias, float* y, bf16* y_hi, bf16* y_lo, int B, int L, float scale, cudaStream_t st);

// ---- rigid.cu -----------------------------------------------------------------------------------------
void frame_update(float* quat, float* trans, const float* upd6, const float* diffuse, int rows, cudaStream_t st);
void bb_update_frame(const float* node, const float* W, const float* bias, float* quat, float* trans, const float* diffuse,
                     int rows, cudaStream_t st);
void split_rigids(const float* rig7, float* quat, float* trans_nm, int rows, cudaStream_t st);
void join_rigids(const float* quat, const float* trans_nm, float* rig7, int rows, cudaStream_t st);
struct Se3StepArgs {
  int B, L;
  const float* rig_t;      // [B,L,7]
  const float* rig_0;      // [B,L,7] predicted clean frames (needed unless scores are given)
  const float* mask;       // [B,L] residue mask (scores are multiplied by it)
  const float* diffuse;    // [B,L] (1-fixed)*mask; nullptr = update everything
  const float* sched_f;    // [B][8]: t, sigma_q, g_rot, g_rot^2, exp(-beta/2), 1-exp(-beta), b_t, sqrt(b_t)
  const double* sched_d;   // [B][2]: dt, sqrt(dt)
  const float* rot_noise;  // [B,L,3] or nullptr (SDE only)
  const float* trans_noise;
  float noise_scale;
  int probability_flow;
  int mode;                // 0 fused score+reverse, 1 score only, 2 reverse from given scores
  double* rot_score;       // [B,L,3] out (mode 0/1, optional) or in (mode 2)
  double* trans_score;
  float* rig_out;          // [B,L,7] (mode 0/2)
};
void se3_step(const Se3StepArgs& a, cudaStream_t st);
struct Se3PerturbArgs {
  int B, L;
  const float* rot0;       // [B,L,9] rotation matrices
  const float* trans0;     // [B,L,3] Angstrom
  const float* diffuse;    // [B,L] or nullptr
  const float* sched_f;    // [B][2]: exp(-beta/2), sqrt(1-exp(-beta))
  const double* cdf;       // [B][1000] IGSO3 cdf row of each decoy's sigma bucket
  const float* omega_grid; // [1000]
  const float* axis_noise; // [B,L,3] N(0,1)
  const float* u_noise;    // [B,L]   U[0,1)
  const float* trans_noise;// [B,L,3] N(0,1)
  float* rig_out;          // [B,L,7]
};
void se3_perturb(const Se3PerturbArgs& a, cudaStream_t st);
void backbone_atoms(const float* rig7, const float* psi, const long long* aatype, const float* table,
                    float* atom37, float* atom14, int rows, cudaStream_t st);

// ---- rng.cu -------------------------------------------------------------------------------------------
void philox_fill(float* out, int B, long n_per_decoy, unsigned long long seed, long long first_decoy, unsigned long long stream_id,
                 int uniform, cudaStream_t st, const long long* decoy_ids = nullptr, const int* stream_ids = nullptr);

}  // namespace s2s
