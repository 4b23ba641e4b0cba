// Row-wise kernels of the node track: LayerNorm(+residual,+mask), key-biased softmax, input features,
// psi head normalisation.  One warp per row; rows are 128..512 floats so everything stays in registers.
#include "row_ops.cuh"
#include "s2s_internal.cuh"

namespace s2s {

namespace {

// y[r] = LN(x[r] (+ res[r])) * w + b, then * rowscale[r]   (one warp per row: layernorm_rows in row_ops.cuh)
// y has pitch D; the optional second copy y2 and the split-bf16 image (y_hi, y_lo) share the pitch ld2, and columns
// [D, D + tail_w) of THOSE rows receive a copy of tail[row][0 .. tail_w) (pitch tail_ld) — the sequence transformer's input
// cat([node, skip_embed(init_node)]) of ipa.py:353-356 written by the LayerNorm that produces node.
template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        const float* __restrict__ rowscale, float* __restrict__ y,
                                                        int rows, bf16* __restrict__ y_hi, bf16* __restrict__ y_lo, int ld2,
                                                        float* __restrict__ y2, const float* __restrict__ tail, int tail_ld, int tail_w) {
  pdl_sync();
  const int row = blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  layernorm_rows<D, 1>(x + (long)row * D, res ? res + (long)row * D : nullptr, D, w, b, rowscale ? rowscale + row : nullptr, y + (long)row * D, D,
                       y_hi ? y_hi + (long)row * ld2 : nullptr, y_lo ? y_lo + (long)row * ld2 : nullptr, y2 ? y2 + (long)row * ld2 : nullptr, ld2, lane, 1);
  if (tail) {
    for (int g = lane; g < tail_w / 4; g += 32) {
      const float4 o = *reinterpret_cast<const float4*>(tail + (long)row * tail_ld + g * 4);
      if (y2) *reinterpret_cast<float4*>(y2 + (long)row * ld2 + D + g * 4) = o;
      if (y_hi) {
        *reinterpret_cast<uint2*>(y_hi + (long)row * ld2 + D + g * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        *reinterpret_cast<uint2*>(y_lo + (long)row * ld2 + D + g * 4) =
            make_uint2(pack_bf16(o.x - bf16_round(o.x), o.y - bf16_round(o.y)), pack_bf16(o.z - bf16_round(o.z), o.w - bf16_round(o.w)));
      }
    }
  }
}

// in-place softmax over the last dim of S[nb][nh][L][L] with an additive per-key bias keybias[b][j]
__global__ void __launch_bounds__(256) softmax_keybias_kernel(float* __restrict__ S, const float* __restrict__ keybias,
                                                              int L, int rows_per_batch, long rows,
                                                              bf16* __restrict__ P_hi, bf16* __restrict__ P_lo) {
  pdl_sync();
  const long row = (long)blockIdx.x * 8 + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= rows) return;
  const int b = (int)(row / rows_per_batch);
  float* p = S + row * L;
  const float* kb = keybias ? keybias + (long)b * L : nullptr;
  if (L <= 32 * 16) {  // the row stays in registers: one read of S, one write of the result
    float v[16];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = lane + 32 * u;
      v[u] = j < L ? p[j] + (kb ? kb[j] : 0.f) : -INFINITY;
      mx = fmaxf(mx, v[u]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      v[u] = __expf(v[u] - mx);
      sum += v[u];
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int j = lane + 32 * u;
      if (j >= L) break;
      const float pv = v[u] * inv;
      if (P_hi) {  // split-bf16 copy for the tensor-core P.V (the fp32 row is then not needed any more)
        const bf16 hi = __float2bfloat16_rn(pv);
        P_hi[row * L + j] = hi;
        if (P_lo) P_lo[row * L + j] = __float2bfloat16_rn(pv - __bfloat162float(hi));
      } else {
        p[j] = pv;
      }
    }
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    float v = p[j] + (kb ? kb[j] : 0.f);
    p[j] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    float e = __expf(p[j] - mx);
    p[j] = e;
    sum += e;
  }
  const float inv = 1.f / warp_sum(sum);
  if (P_hi) {
    for (int j = lane; j < L; j += 32) {
      const float pv = p[j] * inv;
      const bf16 hi = __float2bfloat16_rn(pv);
      P_hi[row * L + j] = hi;
      if (P_lo) P_lo[row * L + j] = __float2bfloat16_rn(pv - __bfloat162float(hi));
    }
    return;
  }
  for (int j = lane; j < L; j += 32) p[j] *= inv;
}

// node features [B*L][65] = [sin(1e4 t f_k) (16) | cos (16) | fixed | sin(idx*pi/d_k) (16) | cos (16)]
// and the 33-wide per-residue time feature tf [B*L][33] used by the pair embedder.
// (reference src/models/net/denoising_ipa.py:13-46,126-146). freq/denom tables come from the host so the
// sin/cos arguments are bit-identical to the reference's fp32 tensors.
__global__ void node_features_kernel(const float* __restrict__ t, const long long* __restrict__ ridx,
                                     const float* __restrict__ fixed, const float* __restrict__ tfreq,
                                     const float* __restrict__ pdenom, float* __restrict__ feat,
                                     float* __restrict__ tf, int B, int L) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * L) return;
  const int b = r / L;
  const float ts = t[b] * 10000.f;
  float* f = feat + (long)r * 65;
  float* g = tf + (long)r * 33;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const float a = ts * tfreq[k];
    const float s = sinf(a), c = cosf(a);
    f[k] = s;
    f[16 + k] = c;
    g[k] = s;
    g[16 + k] = c;
  }
  f[32] = fixed[r];
  g[32] = fixed[r];
  const float p = (float)ridx[r] * 3.14159274101257324f;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const float a = __fdiv_rn(p, pdenom[k]);
    f[33 + k] = sinf(a);
    f[49 + k] = cosf(a);
  }
}

// relative-position feature rows for offsets d_min .. d_min+n-1 : [n][32]
__global__ void relpos_features_kernel(const float* __restrict__ pdenom, float* __restrict__ out, int d_min, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float p = (float)(d_min + r) * 3.14159274101257324f;
  for (int k = 0; k < 16; ++k) {
    const float a = __fdiv_rn(p, pdenom[k]);
    out[(long)r * 32 + k] = sinf(a);
    out[(long)r * 32 + 16 + k] = cosf(a);
  }
}

// psi = u / sqrt(max(|u|^2, 1e-8)); then mixed with the ground-truth psi on fixed residues
// (layers.py:205-213, denoising_ipa.py:192-194)
__global__ void psi_finalize_kernel(const float* __restrict__ u, const float* __restrict__ gt_psi,
                                    const float* __restrict__ fixed, float* __restrict__ psi, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float a = u[2 * r], b = u[2 * r + 1];
  const float inv = 1.f / sqrtf(fmaxf(a * a + b * b, 1e-8f));
  if (!gt_psi) {  // TranslationIPA.forward returns the raw head output (ipa.py:375)
    psi[2 * r] = a * inv;
    psi[2 * r + 1] = b * inv;
    return;
  }
  const float fx = fixed[r];
  psi[2 * r] = gt_psi[2 * r] * fx + a * inv * (1.f - fx);
  psi[2 * r + 1] = gt_psi[2 * r + 1] * fx + b * inv * (1.f - fx);
}

__global__ void concat_skip_kernel(const float* __restrict__ node, const float* __restrict__ skip,
                                   float* __restrict__ out, long rows, bf16* __restrict__ out_hi,
                                   bf16* __restrict__ out_lo) {
  pdl_sync();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * D_TFM) return;
  const long r = i / D_TFM;
  const int c = (int)(i % D_TFM);
  const float v = c < C_S ? node[r * C_S + c] : skip[r * D_SKIP + (c - C_S)];
  out[i] = v;
  if (out_hi) {
    const bf16 hi = __float2bfloat16_rn(v);
    out_hi[i] = hi;
    out_lo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

__global__ void masks_kernel(const float* __restrict__ rmask, const float* __restrict__ fixed, const float* __restrict__ hard,
                             float* __restrict__ diffuse, float* __restrict__ keybias, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  diffuse[i] = (1.f - fixed[i]) * rmask[i];
  // float src_key_padding_mask is ADDED to the logits (ipa.py:357): a masked key of the caller's batch still takes part,
  // with +1.  Rows the library itself appended to reach a tile multiple are not keys at all: exp() of them is exactly 0.
  keybias[i] = (hard && hard[i] == 0.f) ? -1e30f : 1.f - rmask[i];
}

// ---- internal chain-length padding (api.cu: pad_len) ----------------------------------------------------------------
// The tensor-core kernels tile chains in units of 32 residues.  Other lengths run on padded copies [B][Lp] of the
// per-residue inputs: appended rows get mask 0 (IPA keys drop out through the -1e5 mask term, pair rows are zeroed by the
// edge mask), the residue index of the decoy's first residue (offsets stay inside the relative-position table), an identity
// frame, and hard = 0 (dropped from the sequence transformer's keys, see masks_kernel).  Real rows get hard = 1.
__global__ void pad_inputs_kernel(int B, int L, int Lp, const float* __restrict__ rig, const float* __restrict__ sc,
                                  const long long* __restrict__ ridx, const float* __restrict__ rmask,
                                  const float* __restrict__ fixed, const float* __restrict__ psi, float* __restrict__ o_rig,
                                  float* __restrict__ o_sc, long long* __restrict__ o_ridx, float* __restrict__ o_rmask,
                                  float* __restrict__ o_fixed, float* __restrict__ o_psi, float* __restrict__ o_hard) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Lp) return;
  const int b = idx / Lp, j = idx - b * Lp;
  const bool real = j < L;
  const long src = (long)b * L + j;
  if (rig) {
#pragma unroll
    for (int k = 0; k < 7; ++k) o_rig[(long)idx * 7 + k] = real ? rig[src * 7 + k] : (k == 0 ? 1.f : 0.f);
  }
  if (sc) {
#pragma unroll
    for (int k = 0; k < 3; ++k) o_sc[(long)idx * 3 + k] = real ? sc[src * 3 + k] : 0.f;
  }
  if (ridx) o_ridx[idx] = ridx[real ? src : (long)b * L];
  if (rmask) o_rmask[idx] = real ? rmask[src] : 0.f;
  if (fixed) o_fixed[idx] = real ? fixed[src] : 0.f;
  if (psi) {
    o_psi[(long)idx * 2] = real ? psi[src * 2] : 0.f;
    o_psi[(long)idx * 2 + 1] = real ? psi[src * 2 + 1] : 0.f;
  }
  o_hard[idx] = real ? 1.f : 0.f;
}

// dst [B][Ld][W] <- src [B][Ls][W]: rows j < min(Ls, Ld) are copied, rows beyond Ls are filled with zeros
template <typename T>
__global__ void repitch_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int Ls, int Ld, int W) {
  pdl_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * Ld * W) return;
  const int w = (int)(idx % W);
  const long bj = idx / W;
  const int b = (int)(bj / Ld), j = (int)(bj - (long)b * Ld);
  dst[idx] = j < Ls ? src[((long)b * Ls + j) * W + w] : T(0);
}

// pair tensor dst [B][Ld][Ld][128] <- src [B][Ls][Ls][128] (bf16, 16-byte pieces), zero outside the source square
__global__ void repitch_pair_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int Ls, int Ld) {
  pdl_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte piece: 16 per pair row
  if (idx >= (long)B * Ld * Ld * 16) return;
  const int piece = (int)(idx & 15);
  const long row = idx >> 4;
  const int j = (int)(row % Ld);
  const long bi = row / Ld;
  const int i = (int)(bi % Ld), b = (int)(bi / Ld);
  dst[idx] = (i < Ls && j < Ls) ? src[((((long)b * Ls + i) * Ls + j) << 4) + piece] : make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace

void layernorm(const float* x, const float* res, const float* w, const float* b, const float* rowscale, float* y,
               int rows, int D, cudaStream_t st, bf16* y_hi, bf16* y_lo, const LnExtra& ex) {
  S2S_PROF("layernorm", st);
  const int grid = ceil_div(rows, 8);
  const int ld2 = ex.ld2 > 0 ? ex.ld2 : D;
  S2S_CHECK(ld2 % 4 == 0 && ex.tail_w % 4 == 0 && ex.tail_ld % 4 == 0 && (!ex.tail || D + ex.tail_w <= ld2), "layernorm: bad pitch / tail width");
  S2S_CHECK(!y_hi || y_lo, "layernorm: a hi image needs a lo image");
  if (D == 128)
    launch_pdl(layernorm_kernel<128>, grid, 256, 0, st, x, res, w, b, rowscale, y, rows, y_hi, y_lo, ld2, ex.y2, ex.tail, ex.tail_ld, ex.tail_w);
  else if (D == 256)
    launch_pdl(layernorm_kernel<256>, grid, 256, 0, st, x, res, w, b, rowscale, y, rows, y_hi, y_lo, ld2, ex.y2, ex.tail, ex.tail_ld, ex.tail_w);
  else if (D == 320)
    launch_pdl(layernorm_kernel<320>, grid, 256, 0, st, x, res, w, b, rowscale, y, rows, y_hi, y_lo, ld2, ex.y2, ex.tail, ex.tail_ld, ex.tail_w);
  else
    S2S_CHECK(false, "layernorm: unsupported width");
  S2S_LAUNCH_CHECK();
}

void softmax_keybias(float* S, const float* keybias, int nb, int nh, int L, cudaStream_t st, bf16* P_hi, bf16* P_lo) {
  const long rows = (long)nb * nh * L;
  S2S_PROF("softmax", st);
  launch_pdl(softmax_keybias_kernel, ceil_div(rows, 8), 256, 0, st, S, keybias, L, nh * L, rows, P_hi, P_lo);
  S2S_LAUNCH_CHECK();
}

void node_features(const float* t, const long long* ridx, const float* fixed, const float* tfreq,
                   const float* pdenom, float* feat, float* tf, int B, int L, cudaStream_t st) {
  launch_pdl(node_features_kernel, ceil_div((long)B * L, 128), 128, 0, st, t, ridx, fixed, tfreq, pdenom, feat, tf, B, L);
  S2S_LAUNCH_CHECK();
}

void relpos_features(const float* pdenom, float* out, int d_min, int n, cudaStream_t st) {
  relpos_features_kernel<<<ceil_div(n, 128), 128, 0, st>>>(pdenom, out, d_min, n);
  S2S_LAUNCH_CHECK();
}

void psi_finalize(const float* u, const float* gt_psi, const float* fixed, float* psi, int rows, cudaStream_t st) {
  launch_pdl(psi_finalize_kernel, ceil_div(rows, 128), 128, 0, st, u, gt_psi, fixed, psi, rows);
  S2S_LAUNCH_CHECK();
}

void concat_skip(const float* node, const float* skip, float* out, long rows, cudaStream_t st, bf16* out_hi, bf16* out_lo) {
  S2S_PROF("concat_skip", st);
  launch_pdl(concat_skip_kernel, ceil_div(rows * D_TFM, 256), 256, 0, st, node, skip, out, rows, out_hi, out_lo);
  S2S_LAUNCH_CHECK();
}

void make_masks(const float* rmask, const float* fixed, const float* hard, float* diffuse, float* keybias, int n, cudaStream_t st) {
  launch_pdl(masks_kernel, ceil_div(n, 256), 256, 0, st, rmask, fixed, hard, diffuse, keybias, n);
  S2S_LAUNCH_CHECK();
}

void pad_inputs(const PadInputs& a, cudaStream_t st) {
  launch_pdl(pad_inputs_kernel, ceil_div((long)a.B * a.Lp, 128), 128, 0, st, a.B, a.L, a.Lp, a.rig, a.sc, a.ridx, a.rmask, a.fixed, a.psi, a.o_rig,
                                                                     a.o_sc, a.o_ridx, a.o_rmask, a.o_fixed, a.o_psi, a.o_hard);
  S2S_LAUNCH_CHECK();
}
void repitch_rows(const float* src, float* dst, int B, int Ls, int Ld, int W, cudaStream_t st) {
  launch_pdl(repitch_rows_kernel<float>, ceil_div((long)B * Ld * W, 256), 256, 0, st, src, dst, B, Ls, Ld, W);
  S2S_LAUNCH_CHECK();
}
void repitch_pair(const bf16* src, bf16* dst, int B, int Ls, int Ld, cudaStream_t st) {
  launch_pdl(repitch_pair_kernel, ceil_div((long)B * Ld * Ld * 16, 256), 256, 0, st, reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), B, Ls, Ld);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
