// SIMT restatement of the two pair-track MLP kernels (edge embedder, EdgeTransition).
//
// Same inputs, outputs and bf16 rounding points as the tcgen05 kernels (pair_tc3.cu, pair_tc4.cu), but with plain FFMA
// loops: the on-device cross-check for the tensor-core path (pair_kernels = 0; tests/test_gpu_parity.py).  The product
// path never takes it: chain lengths are padded to the tensor-core tiling inside the library (api.cu).  Not a CPU fallback.
#include "s2s_internal.cuh"

namespace s2s {

namespace {
constexpr int RT = 32;  // pair rows per CTA

__device__ __forceinline__ void row_to_bij(long r, int L, int& b, int& i, int& j) {
  j = (int)(r % L);
  const long t = r / L;
  i = (int)(t % L);
  b = (int)(t / L);
}

// out[n] (for RT rows) += sum_k Wt[k][n] * xT[k][0..RT)
template <int K>
__device__ __forceinline__ void accum_rows(float (&acc)[RT], const bf16* __restrict__ Wt, int ldw, int n,
                                           const float* __restrict__ xT) {
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float w = __bfloat162float(Wt[(long)k * ldw + n]);
    const float4* xr = reinterpret_cast<const float4*>(xT + k * RT);
#pragma unroll
    for (int r4 = 0; r4 < RT / 4; ++r4) {
      const float4 x = xr[r4];
      acc[r4 * 4 + 0] = fmaf(w, x.x, acc[r4 * 4 + 0]);
      acc[r4 * 4 + 1] = fmaf(w, x.y, acc[r4 * 4 + 1]);
      acc[r4 * 4 + 2] = fmaf(w, x.z, acc[r4 * 4 + 2]);
      acc[r4 * 4 + 3] = fmaf(w, x.w, acc[r4 * 4 + 3]);
    }
  }
}

// LayerNorm(128) of RT rows held in y_s[r][128], times rowmask, stored as bf16
__device__ __forceinline__ void ln_store_rows(const float* __restrict__ y_s, const float* __restrict__ lw,
                                              const float* __restrict__ lb, const float* __restrict__ mask, int L,
                                              long row0, long rows, bf16* __restrict__ out, int nwarps) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int r = warp; r < RT; r += nwarps) {
    const long row = row0 + r;
    if (row >= rows) continue;
    float v[4];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = y_s[r * C_Z + lane * 4 + u];
      s += v[u];
    }
    const float mean = warp_sum(s) * (1.f / C_Z);
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) q += (v[u] - mean) * (v[u] - mean);
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C_Z) + 1e-5f);
    int b, i, j;
    row_to_bij(row, L, b, i, j);
    const float m = mask[(long)b * L + i] * mask[(long)b * L + j];
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = lane * 4 + u;
      o[u] = ((v[u] - mean) * rstd * lw[c] + lb[c]) * m;
    }
    uint2 pk = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    *reinterpret_cast<uint2*>(out + row * C_Z + lane * 4) = pk;
  }
}

// ---- edge embedder ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) edge_embed_simt_kernel(EdgeEmbedArgs a) {
  pdl_sync();
  extern __shared__ float smem[];
  float* h0T = smem;                 // [128][RT]
  float* h1T = smem + C_Z * RT;      // [128][RT]
  float* y_s = h0T;                  // reused after layer 2: [RT][128]
  const int n = threadIdx.x;
  const long rows = (long)a.B * a.L * a.L;
  const long row0 = (long)blockIdx.x * RT;

  // layer 1 as table lookups (denoising_ipa.py:126-158)
  for (int r = 0; r < RT; ++r) {
    const long row = row0 + r;
    float h = 0.f;
    if (row < rows) {
      int b, i, j;
      row_to_bij(row, a.L, b, i, j);
      const long bi = (long)b * a.L + i, bj = (long)b * a.L + j;
      const int bin = pair_distogram_bin(a.sc_ca + bi * 3, a.sc_ca + bj * 3, a.bin_lower);
      const int off = min(max((int)(a.ridx[bi] - a.ridx[bj]) - a.d_min, 0), a.n_off - 1);
      h = a.Ti[bi * C_Z + n] + a.Tj[bj * C_Z + n] + a.Tpos[(long)off * C_Z + n];
      if (bin >= 0) h += a.Wd[bin * C_Z + n];
      h = bf16_round(fmaxf(h, 0.f));
    }
    h0T[n * RT + r] = h;
  }
  __syncthreads();
  float acc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
  accum_rows<C_Z>(acc, a.W2t, C_Z, n, h0T);
#pragma unroll
  for (int r = 0; r < RT; ++r) h1T[n * RT + r] = bf16_round(fmaxf(acc[r] + a.b2[n], 0.f));
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
  accum_rows<C_Z>(acc, a.W3t, C_Z, n, h1T);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RT; ++r) y_s[r * C_Z + n] = acc[r] + a.b3[n];
  __syncthreads();
  ln_store_rows(y_s, a.ln_w, a.ln_b, a.mask, a.L, row0, rows, a.z_out, 4);
}

// ---- EdgeTransition -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(D_ET) edge_transition_simt_kernel(EdgeTransitionArgs a) {
  pdl_sync();
  extern __shared__ float smem[];
  float* xT = smem;                          // [128][RT]   z rows (bf16 values)
  float* h1T = xT + C_Z * RT;                // [384][RT]
  float* h2T = h1T + D_ET * RT;              // [384][RT]
  float* y_s = h1T;                          // reused: [RT][128]
  const int n = threadIdx.x;
  const long rows = (long)a.B * a.L * a.L;
  const long row0 = (long)blockIdx.x * RT;

  for (int idx = n; idx < RT * C_Z; idx += D_ET) {
    const int r = idx / C_Z, c = idx % C_Z;
    const long row = row0 + r;
    xT[c * RT + r] = row < rows ? __bfloat162float(a.z_in[row * C_Z + c]) : 0.f;
  }
  __syncthreads();
  float acc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
  accum_rows<C_Z>(acc, a.W1zt, D_ET, n, xT);
  for (int r = 0; r < RT; ++r) {
    const long row = row0 + r;
    float h = 0.f;
    if (row < rows) {
      int b, i, j;
      row_to_bij(row, a.L, b, i, j);
      h = acc[r] + a.u[((long)b * a.L + i) * D_ET + n] + a.v[((long)b * a.L + j) * D_ET + n];
    }
    h1T[n * RT + r] = bf16_round(fmaxf(h, 0.f));
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
  accum_rows<D_ET>(acc, a.W2t, D_ET, n, h1T);
#pragma unroll
  for (int r = 0; r < RT; ++r) h2T[n * RT + r] = bf16_round(fmaxf(acc[r] + a.b2[n], 0.f));
  __syncthreads();
  if (n < C_Z) {
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.f;
    accum_rows<D_ET>(acc, a.Wft, C_Z, n, h2T);
    accum_rows<C_Z>(acc, a.Wfzt, C_Z, n, xT);
  }
  __syncthreads();  // all reads of h1T done before y_s (aliases h1T) is written
  if (n < C_Z) {
    for (int r = 0; r < RT; ++r) {
      const long row = row0 + r;
      float y = 0.f;
      if (row < rows) {
        int b, i, j;
        row_to_bij(row, a.L, b, i, j);
        y = acc[r] + a.p[((long)b * a.L + i) * C_Z + n] + a.q[((long)b * a.L + j) * C_Z + n];
      }
      y_s[r * C_Z + n] = y;
    }
  }
  __syncthreads();
  ln_store_rows(y_s, a.ln_w, a.ln_b, a.mask, a.L, row0, rows, a.z_out, D_ET / 32);
}

}  // namespace

void edge_embed_simt(const EdgeEmbedArgs& a, cudaStream_t st) {
  const long rows = (long)a.B * a.L * a.L;
  const size_t smem = 2 * C_Z * RT * sizeof(float);
  S2S_PROF("edge_embed_simt", st);
  launch_pdl(edge_embed_simt_kernel, ceil_div(rows, RT), 128, smem, st, a);
  S2S_LAUNCH_CHECK();
}

void edge_transition_simt(const EdgeTransitionArgs& a, cudaStream_t st) {
  const long rows = (long)a.B * a.L * a.L;
  const size_t smem = (C_Z + 2 * D_ET) * RT * sizeof(float);
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  S2S_PROF("edge_transition_simt", st);
  launch_pdl(edge_transition_simt_kernel, ceil_div(rows, RT), D_ET, smem, st, a);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
