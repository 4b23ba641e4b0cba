// Invariant Point Attention kernels (reference src/models/net/ipa.py:100-268).
//
// The pair tensor z[b,i,:,:] (L x 128 bf16) is read from HBM exactly once per IPA call, by
// ipa_pair_attention_kernel: one CTA per (decoy, query residue) stages the slab in shared memory and uses it
// for BOTH the pair bias (linear_b, before the softmax) and the pair value aggregation (down_z, after it).
// The q.k logits and the point term arrive in S (written by a batched GEMM + ipa_point_logits); the
// attention weights go back into S for the P*V / P*V_pts products.
#include "s2s_internal.cuh"

namespace s2s {

namespace {

// raw point projections -> global-frame points (ipa.py:144-171): x|y|z chunk layout, R p + t (t in nm).
// With `aug` set the kernel also writes the tensor-core operands of the fused logits / value GEMMs:
//   qp_aug / kp_aug [row][head][PT_K] bf16: sqrt(w_h) * points as (hi | lo | hi) and (hi | hi | lo) split-bf16 triples, so that
//     one bf16 MMA over these 72 (+8 zero) columns gives  w_h q_pts.k_pts  to ~2^-17 relative;
//   colbias [b][h][j] = -0.5 w_h |k_pts_j|^2.  Together: -0.5 w_h |q - k|^2 up to a per-query constant, which the
//     softmax over keys cancels exactly (ipa.py:191-205,215);
//   vp_hi / vp_lo [row][8][40]: split-bf16 value points (36 + 4 zero columns per head), row-major (read as the MN-major B operand of the P.v_pts GEMM).
// In this mode the fp32 point arrays are not written at all.
__global__ void __launch_bounds__(256) ipa_points_kernel(const float* __restrict__ qp, long ld_q,
                                                         const float* __restrict__ kvp, long ld_kv,
                                                         const float* __restrict__ quat,
                                                         const float* __restrict__ trans, float* __restrict__ q_pts,
                                                         float* __restrict__ k_pts, float* __restrict__ v_pts,
                                                         int rows, IpaPointsAug aug) {
  pdl_sync();
  __shared__ float k2_s[N_H * P_Q];
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  if (r >= rows) return;
  float q[4] = {quat[r * 4], quat[r * 4 + 1], quat[r * 4 + 2], quat[r * 4 + 3]};
  float R[9];
  quat_to_rot(q, R);
  const float tx = trans[r * 3], ty = trans[r * 3 + 1], tz = trans[r * 3 + 2];
  constexpr int NQ = N_H * P_Q, NKV = N_H * (P_Q + P_V);
  const bool active = tid < NQ + NKV;
  int kind = -1, h = 0, p = 0;  // 0 query point, 1 key point, 2 value point
  float o[3] = {0.f, 0.f, 0.f};
  if (active) {
    float x, y, z;
    float* dst;
    if (tid < NQ) {
      x = qp[r * ld_q + tid];
      y = qp[r * ld_q + NQ + tid];
      z = qp[r * ld_q + 2 * NQ + tid];
      dst = q_pts + ((long)r * NQ + tid) * 3;
      kind = 0; h = tid / P_Q; p = tid % P_Q;
    } else {
      const int k = tid - NQ;
      h = k / (P_Q + P_V); p = k % (P_Q + P_V);
      x = kvp[r * ld_kv + k];
      y = kvp[r * ld_kv + NKV + k];
      z = kvp[r * ld_kv + 2 * NKV + k];
      if (p < P_Q) { kind = 1; dst = k_pts + (((long)r * N_H + h) * P_Q + p) * 3; }
      else { kind = 2; p -= P_Q; dst = v_pts + (((long)r * N_H + h) * P_V + p) * 3; }
    }
    o[0] = R[0] * x + R[1] * y + R[2] * z + tx;
    o[1] = R[3] * x + R[4] * y + R[5] * z + ty;
    o[2] = R[6] * x + R[7] * y + R[8] * z + tz;
    if (!aug.qp_aug) { dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; }  // fp32 points: first-generation / exact path only
  }
  if (!aug.qp_aug) return;  // uniform
  // tensor-core operands: built in shared memory, written out as whole 16-byte chunks (one per thread)
  __shared__ __align__(16) bf16 row_s[2 * N_H * PT_K + 2 * N_H * VP_PITCH];  // q row | k row | v hi | v lo
  bf16* q_row = row_s;
  bf16* k_row = row_s + N_H * PT_K;
  bf16* v_hi = row_s + 2 * N_H * PT_K;
  bf16* v_lo = v_hi + N_H * VP_PITCH;
  const int b = r / aug.L, j = r % aug.L;
  if (kind == 0 || kind == 1) {
    const float w = aug.pt_w[h];
    const float sw = sqrtf(w * aug.inv_alpha);  // the GEMM epilogue multiplies the whole accumulator by alpha
    bf16* dst = (kind == 0 ? q_row : k_row) + h * PT_K + p * 3;
    float k2 = 0.f;
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const float v = o[e] * sw;
      const bf16 hi = __float2bfloat16_rn(v);
      const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      dst[e] = hi;
      dst[24 + e] = kind == 0 ? lo : hi;
      dst[48 + e] = kind == 0 ? hi : lo;
      k2 += o[e] * o[e] * w;
    }
    if (p == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[72 + e] = __float2bfloat16_rn(0.f);
    }
    if (kind == 1) k2_s[h * P_Q + p] = k2;
  } else if (kind == 2) {
#pragma unroll
    for (int e = 0; e < 3; ++e) {
      const bf16 hi = __float2bfloat16_rn(o[e]);
      v_hi[h * VP_PITCH + p * 3 + e] = hi;
      v_lo[h * VP_PITCH + p * 3 + e] = __float2bfloat16_rn(o[e] - __bfloat162float(hi));
    }
    if (p == 0) {  // the 4 pad columns of this head (a TMA box must start on a 16-byte boundary: 40-column head pitch)
#pragma unroll
      for (int e = P_V * 3; e < VP_PITCH; ++e) v_hi[h * VP_PITCH + e] = v_lo[h * VP_PITCH + e] = __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  {
    constexpr int NQC = N_H * PT_K / 8, NVC = N_H * VP_PITCH / 8;  // 16-byte chunks per row: 80 and 40
    const uint4* src = reinterpret_cast<const uint4*>(row_s);
    if (tid < NQC) reinterpret_cast<uint4*>(aug.qp_aug + (long)r * N_H * PT_K)[tid] = src[tid];
    else if (tid < 2 * NQC) reinterpret_cast<uint4*>(aug.kp_aug + (long)r * N_H * PT_K)[tid - NQC] = src[tid];
    else if (tid < 2 * NQC + NVC) reinterpret_cast<uint4*>(aug.vp_hi + (long)r * N_H * VP_PITCH)[tid - 2 * NQC] = src[tid];
    else if (tid < 2 * NQC + 2 * NVC) reinterpret_cast<uint4*>(aug.vp_lo + (long)r * N_H * VP_PITCH)[tid - 2 * NQC - NVC] = src[tid];
  }
  if (tid < N_H) {
    float s = 0.f;
#pragma unroll
    for (int pp = 0; pp < P_Q; ++pp) s += k2_s[tid * P_Q + pp];
    aug.colbias[((long)b * N_H + tid) * aug.L + j] = -0.5f * s;
  }
}

// S[b,h,i,j] += -0.5 * w_h * sum_p |q_pts[b,i,h,p] - k_pts[b,j,h,p]|^2      (ipa.py:191-205)
// Direct differences in fp32: the |q|^2+|k|^2-2q.k expansion cancels catastrophically for nm-scale coordinates.
// Block = 32 queries x 64 keys of one (decoy, head); a thread keeps one key's 8 points in registers and walks 8 queries.
__global__ void __launch_bounds__(256) ipa_point_logits_kernel(float* __restrict__ S, const float* __restrict__ q_pts,
                                                               const float* __restrict__ k_pts,
                                                               const float* __restrict__ pt_w, int L) {
  pdl_sync();
  __shared__ __align__(16) float qs[32][P_Q * 3];
  const int bh = blockIdx.z, b = bh / N_H, h = bh % N_H;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 32 * 24; idx += 256) {
    const int il = idx / 24, c = idx % 24;
    qs[il][c] = (i0 + il < L) ? q_pts[(((long)b * L + i0 + il) * N_H + h) * 24 + c] : 0.f;
  }
  const int jl = tid % 64, ig = tid / 64;
  const int j = j0 + jl;
  float kp[24];
  if (j < L) {
    const float4* src = reinterpret_cast<const float4*>(k_pts + (((long)b * L + j) * N_H + h) * 24);
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const float4 v = __ldg(src + u);
      kp[4 * u] = v.x; kp[4 * u + 1] = v.y; kp[4 * u + 2] = v.z; kp[4 * u + 3] = v.w;
    }
  }
  __syncthreads();
  if (j >= L) return;
  const float w = pt_w[h];
  float* Sp = S + (((long)b * N_H + h) * L + i0 + ig * 8) * L + j;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int il = ig * 8 + u;
    if (i0 + il >= L) break;
    float q[24];
#pragma unroll
    for (int v4 = 0; v4 < 6; ++v4) {  // warp-uniform address: one broadcast 128-bit read per 4 coordinates
      const float4 t = *reinterpret_cast<const float4*>(&qs[il][4 * v4]);
      q[4 * v4] = t.x; q[4 * v4 + 1] = t.y; q[4 * v4 + 2] = t.z; q[4 * v4 + 3] = t.w;
    }
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < P_Q; ++p) {
      const float dx = q[p * 3] - kp[p * 3], dy = q[p * 3 + 1] - kp[p * 3 + 1], dz = q[p * 3 + 2] - kp[p * 3 + 2];
      acc += (dx * dx + dy * dy + dz * dz) * w;
    }
    Sp[(long)u * L] += acc * (-0.5f);
  }
}

// ---- the pair kernel ----------------------------------------------------------------------------------
constexpr int ZP = C_Z + 8;  // padded bf16 row pitch (272 B): conflict-free ldmatrix

__global__ void __launch_bounds__(256) ipa_pair_attention_kernel(IpaPairArgs a) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = a.L, Lp = (L + 15) & ~15;
  const int LP4 = Lp + 4, LP8 = Lp + 8;
  bf16* z_s = reinterpret_cast<bf16*>(smem_raw);                      // [Lp][ZP]
  float* P_s = reinterpret_cast<float*>(z_s + (size_t)Lp * ZP);        // [8][LP4] logits -> probabilities
  bf16* Ph = reinterpret_cast<bf16*>(P_s + 8 * LP4);                   // [8][LP8] bf16 hi part of P
  bf16* Pl = Ph + 8 * LP8;                                             // [8][LP8] bf16 lo part of P
  float* zsum = reinterpret_cast<float*>(Pl + 8 * LP8);                // [8][C_Z+4]  sum_j P[h][j] z[j][:]
  float* mask_s = zsum + 8 * (C_Z + 4);                                // [Lp] key mask of this decoy
  float* wdz_s = mask_s + Lp;                                          // [128][32] down_z weight (transposed)

  const int b = blockIdx.x / L, i = blockIdx.x % L;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int g = lane / 4, t = lane % 4;

  // ---- stage the z slab (one HBM read of z per IPA call) and the logits row block ----------------------
  const bf16* zg = a.z + ((long)b * L + i) * L * C_Z;
  for (int idx = tid; idx < L * (C_Z / 8); idx += 256) {
    const int j = idx / (C_Z / 8), c8 = idx % (C_Z / 8);
    cp_async16(z_s + j * ZP + c8 * 8, zg + (long)j * C_Z + c8 * 8);
  }
  for (int idx = tid; idx < C_Z * 32 / 4; idx += 256) cp_async16(wdz_s + idx * 4, a.Wdz_t + idx * 4);
  cp_async_commit();
  for (int j = tid; j < Lp; j += 256) mask_s[j] = j < L ? a.mask[(long)b * L + j] : 0.f;
  for (int idx = tid; idx < (Lp - L) * (C_Z / 8); idx += 256) {
    const int j = L + idx / (C_Z / 8), c8 = idx % (C_Z / 8);
    *reinterpret_cast<uint4*>(z_s + j * ZP + c8 * 8) = make_uint4(0, 0, 0, 0);
  }
  {
    const float* Srow = a.S + (((long)b * N_H + warp) * L + i) * L;
    for (int j = lane; j < Lp; j += 32) P_s[warp * LP4 + j] = j < L ? Srow[j] : 0.f;
  }
  // linear_b weight fragments (B operand: k = channel, n = head), hi and lo bf16 terms
  uint32_t wbh[8][2], wbl[8][2];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int k0 = ks * 16 + 2 * t;
    wbh[ks][0] = *reinterpret_cast<const uint32_t*>(a.Wb_hi + g * C_Z + k0);
    wbh[ks][1] = *reinterpret_cast<const uint32_t*>(a.Wb_hi + g * C_Z + k0 + 8);
    wbl[ks][0] = *reinterpret_cast<const uint32_t*>(a.Wb_lo + g * C_Z + k0);
    wbl[ks][1] = *reinterpret_cast<const uint32_t*>(a.Wb_lo + g * C_Z + k0 + 8);
  }
  const float m_i = a.mask[(long)b * L + i];
  const float bias_h[2] = {a.bb[2 * t], a.bb[2 * t + 1]};  // this lane's two heads in the bias MMA fragment
  cp_async_wait<0>();
  __syncthreads();

  // ---- pair bias: [16 keys x 128] x [128 x 8 heads] per m-tile, added into the logits ------------------
  constexpr float SQRT1_3 = 0.57735026918962576f;
  for (int mt = warp; mt < Lp / 16; mt += 8) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* arow = z_s + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ZP + (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t af[4];
      ldmatrix_x4(af, arow + ks * 16);
      mma_bf16_16816(d, af, wbh[ks]);
      mma_bf16_16816(d, af, wbl[ks]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int h = 2 * t + (e & 1);
      const int j = mt * 16 + g + (e >> 1) * 8;
      if (j < L) P_s[h * LP4 + j] += SQRT1_3 * (d[e] + bias_h[e & 1]) + 1e5f * (m_i * mask_s[j] - 1.f);
    }
  }
  __syncthreads();

  // ---- softmax over keys, one warp per head -------------------------------------------------------------
  {
    float* row = P_s + warp * LP4;
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
    float* Srow = a.S + (((long)b * N_H + warp) * L + i) * L;
    for (int j = lane; j < Lp; j += 32) {
      const float pv = j < L ? row[j] * inv : 0.f;
      if (j < L) Srow[j] = pv;
      const bf16 hi = __float2bfloat16_rn(pv);
      if (a.P_bf16 && j < L) a.P_bf16[(((long)b * N_H + warp) * L + i) * L + j] = hi;
      Ph[warp * LP8 + j] = hi;
      Pl[warp * LP8 + j] = __float2bfloat16_rn(pv - __bfloat162float(hi));
    }
  }
  __syncthreads();

  // ---- zsum^T[c][h] = sum_j z[j][c] P[h][j] : warp w owns channels 16w..16w+15 --------------------------
  {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    const bf16* arow = z_s + ((lane >> 4) * 8 + (lane & 7)) * ZP + warp * 16 + ((lane >> 3) & 1) * 8;
    const bf16* ph = Ph + g * LP8 + 2 * t;
    const bf16* pl = Pl + g * LP8 + 2 * t;
    for (int ks = 0; ks < Lp / 16; ++ks) {
      uint32_t af[4], bh[2], bl[2];
      ldmatrix_x4_trans(af, arow + (size_t)ks * 16 * ZP);
      bh[0] = *reinterpret_cast<const uint32_t*>(ph + ks * 16);
      bh[1] = *reinterpret_cast<const uint32_t*>(ph + ks * 16 + 8);
      bl[0] = *reinterpret_cast<const uint32_t*>(pl + ks * 16);
      bl[1] = *reinterpret_cast<const uint32_t*>(pl + ks * 16 + 8);
      mma_bf16_16816(d, af, bh);
      mma_bf16_16816(d, af, bl);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int h = 2 * t + (e & 1);
      const int c = warp * 16 + g + (e >> 1) * 8;
      zsum[h * (C_Z + 4) + c] = d[e];
    }
  }
  __syncthreads();

  // ---- o_pair[h][d] = down_z(zsum[h]) (sum_j P = 1, so the bias passes through; ipa.py:253-254) ---------
  {
    const int h = tid / 32, dd = tid % 32;
    float acc = a.bdz[dd];
    const float* zs = zsum + h * (C_Z + 4);
#pragma unroll 8
    for (int c = 0; c < C_Z; ++c) acc = fmaf(wdz_s[c * 32 + dd], zs[c], acc);
    a.o_pair[((long)b * L + i) * a.ld_opair + h * 32 + dd] = acc;
  }
}

// o_pt (global frame, [rows][H][P_V][3]) -> local frame, norms -> feature columns (ipa.py:229-248,259)
__global__ void __launch_bounds__(96) ipa_finalize_points_kernel(const float* __restrict__ opt,
                                                                  const float* __restrict__ quat,
                                                                  const float* __restrict__ trans,
                                                                  float* __restrict__ feats, int rows,
                                                                  bf16* __restrict__ f_hi, bf16* __restrict__ f_lo) {
  pdl_sync();
  const int r = blockIdx.x, k = threadIdx.x;  // k = h*12 + p
  if (r >= rows) return;
  float q[4] = {quat[r * 4], quat[r * 4 + 1], quat[r * 4 + 2], quat[r * 4 + 3]};
  float R[9];
  quat_to_rot(q, R);
  const float* o = opt + ((long)r * N_H * P_V + k) * 3;
  const float x = o[0] - trans[r * 3], y = o[1] - trans[r * 3 + 1], z = o[2] - trans[r * 3 + 2];
  const float lx = R[0] * x + R[3] * y + R[6] * z;
  const float ly = R[1] * x + R[4] * y + R[7] * z;
  const float lz = R[2] * x + R[5] * y + R[8] * z;
  constexpr int NP = N_H * P_V;
  const float ov[4] = {lx, ly, lz, sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f)};
  const long base = (long)r * IPA_FEAT + N_H * C_H + k;
  if (f_hi) {  // split-bf16 image for linear_out on the tensor cores (no fp32 copy needed)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bf16 h = __float2bfloat16_rn(ov[e]);
      f_hi[base + e * NP] = h;
      f_lo[base + e * NP] = __float2bfloat16_rn(ov[e] - __bfloat162float(h));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) feats[base + e * NP] = ov[e];
  }
}

__global__ void softplus_point_weights_kernel(const float* __restrict__ hw, float* __restrict__ out) {
  const int h = threadIdx.x;
  if (h >= N_H) return;
  const float x = hw[h];
  const float sp = x > 20.f ? x : log1pf(expf(x));  // torch.nn.Softplus(beta=1, threshold=20)
  out[h] = sp * 0.09622504486493763f;               // sqrt(1 / (3 * (8 * 9 / 2)))  (ipa.py:198-200)
}

}  // namespace

void ipa_points(const float* qp_raw, long ld_q, const float* kvp_raw, long ld_kv, const float* quat,
                const float* trans, float* q_pts, float* k_pts, float* v_pts, int rows, cudaStream_t st,
                const IpaPointsAug& aug) {
  S2S_PROF("ipa_points", st);
  S2S_CHECK(!aug.qp_aug || (aug.kp_aug && aug.colbias && aug.vp_hi && aug.vp_lo && aug.pt_w && aug.L > 0), "ipa_points: incomplete operand set");
  launch_pdl(ipa_points_kernel, rows, 256, 0, st, qp_raw, ld_q, kvp_raw, ld_kv, quat, trans, q_pts, k_pts, v_pts, rows, aug);
  S2S_LAUNCH_CHECK();

}

void ipa_point_logits(float* S, const float* q_pts, const float* k_pts, const float* pt_w, int B, int L,
                      cudaStream_t st) {
  S2S_PROF("ipa_point_logits", st);
  dim3 grid(ceil_div(L, 64), ceil_div(L, 32), B * N_H);
  launch_pdl(ipa_point_logits_kernel, grid, 256, 0, st, S, q_pts, k_pts, pt_w, L);
  S2S_LAUNCH_CHECK();
}

void ipa_pair_attention(const IpaPairArgs& a, cudaStream_t st) {
  const int Lp = (a.L + 15) & ~15;
  const size_t smem = (size_t)Lp * ZP * 2 + 8 * (Lp + 4) * 4 + 2 * 8 * (Lp + 8) * 2 + 8 * (C_Z + 4) * 4 + Lp * 4 + C_Z * 32 * 4;
  S2S_CHECK(smem <= 227 * 1024, "ipa_pair_attention: chain too long for one shared-memory slab (L <= 704)");
  static size_t configured = 0;
  if (smem > configured) {
    S2S_CUDA(cudaFuncSetAttribute(ipa_pair_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  S2S_PROF("ipa_pair_attention", st);
  launch_pdl(ipa_pair_attention_kernel, a.B * a.L, 256, smem, st, a);
  S2S_LAUNCH_CHECK();
}

void ipa_finalize_points(const float* opt_glob, const float* quat, const float* trans, float* feats, int rows,
                         cudaStream_t st, bf16* f_hi, bf16* f_lo) {
  S2S_PROF("ipa_finalize_points", st);
  launch_pdl(ipa_finalize_points_kernel, rows, 96, 0, st, opt_glob, quat, trans, feats, rows, f_hi, f_lo);
  S2S_LAUNCH_CHECK();
}

void softplus_point_weights(const float* head_w, float* pt_w, cudaStream_t st) {
  softplus_point_weights_kernel<<<1, 32, 0, st>>>(head_w, pt_w);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
