// Exact-fp32 batched GEMM with a fused epilogue (SIMT FFMA).
//
// Used where the parity budget needs true fp32 products (bb_update, node_embed: tools/precision_probe.py)
// and as the always-correct path for every other node-side linear / attention product.
//   C[b,h][m][n] = post( relu?( alpha * sum_k A[m][k] * B(n,k) * row_pre[m] + bias[n] ) ) * row_post[m] + res[m][n]
// B is "weight-like" [N][K] (b_kn = 0, nn.Linear layout) or [K][N] (b_kn = 1, e.g. P*V).
#include "s2s_internal.cuh"

namespace s2s {

namespace {
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, NT = 256;

template <bool B_KN>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(GemmArgs g) {
  pdl_sync();
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int bz = blockIdx.z;
  const int ib = bz / g.nh, ih = bz % g.nh;
  const float* __restrict__ A = g.A + ib * g.sAb + ih * g.sAh;
  const float* __restrict__ Bm = g.B + ib * g.sBb + ih * g.sBh;
  float* __restrict__ C = g.C + ib * g.sCb + ih * g.sCh;
  const float* __restrict__ R = g.res ? g.res + ib * g.sCb + ih * g.sCh : nullptr;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each a 4x4 micro-tile

  // loader mapping: A tile 64 rows x 16 k -> thread loads 4 consecutive k of one row
  const int a_row = tid / 4, a_k = (tid % 4) * 4;
  // B [N][K]: same mapping; B [K][N]: thread loads 4 consecutive n of one k
  const int b_row = B_KN ? tid / 16 : tid / 4, b_col = B_KN ? (tid % 16) * 4 : (tid % 4) * 4;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto load_tiles = [&](int k0) {
    const int m = m0 + a_row;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + a_k + u;
      ra[u] = (m < g.M && k < g.K) ? A[(long)m * g.lda + k] : 0.f;
    }
    if (B_KN) {
      const int k = k0 + b_row;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + b_col + u;
        rb[u] = (k < g.K && n < g.N) ? Bm[(long)k * g.ldb + n] : 0.f;
      }
    } else {
      const int n = n0 + b_row;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + b_col + u;
        rb[u] = (n < g.N && k < g.K) ? Bm[(long)n * g.ldb + k] : 0.f;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 4; ++u) As[buf][a_k + u][a_row] = ra[u];
    if (B_KN) {
#pragma unroll
      for (int u = 0; u < 4; ++u) Bs[buf][b_row][b_col + u] = rb[u];
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) Bs[buf][b_col + u][b_row] = rb[u];
    }
  };

  const int nk = (g.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
    const float pre = g.row_pre ? g.row_pre[m] : 1.f;
    const float post = g.row_post ? g.row_post[m] : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha * pre;
      if (g.bias) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      v *= post;
      if (R) v += R[(long)m * g.ldres + n];
      C[(long)m * g.ldc + n] = v;
    }
  }
}
}  // namespace

void gemm_f32(const GemmArgs& g, cudaStream_t st) {
  S2S_CHECK(g.M > 0 && g.N > 0 && g.K > 0 && g.nb > 0 && g.nh > 0, "gemm: bad shape");
  S2S_PROF(g_profile_on ? prof_intern("gemm_f32 M" + std::to_string(g.M) + " N" + std::to_string(g.N) + " K" + std::to_string(g.K) + " b" + std::to_string(g.nb * g.nh)) : "gemm", st);
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), g.nb * g.nh);
  if (g.b_kn)
    launch_pdl(gemm_f32_kernel<true>, grid, NT, 0, st, g);
  else
    launch_pdl(gemm_f32_kernel<false>, grid, NT, 0, st, g);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
