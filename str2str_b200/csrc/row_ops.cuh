// Per-row device functions shared by the row kernels (rows.cu) and the fused GEMM chain (gemm_chain.cu).
#pragma once
#include "common.cuh"

namespace s2s {

// One warp normalises NR rows of D channels:  y = (LN(x (+ res)) * w + b) * sc, optionally with its split-bf16 image and a
// second fp32 copy y2.  A lane owns groups of 4 consecutive channels (16-byte loads and stores, 8-byte bf16 image stores).
// The loads of all NR rows are issued before the first reduction, so a warp that walks many rows (the LayerNorm epilogue of
// gemm_chain.cu) exposes one memory latency per NR rows instead of one per row; the arithmetic of a row does not depend on NR.
// Row i lives at x + i * ldx (res + i * ldx when given), y + i * ldy; y2 and the image share the pitch ld2; sc[i] scales row i
// (nullptr: 1); rows i >= n_valid are skipped.  y may alias x.
template <int D, int NR>
__device__ __forceinline__ void layernorm_rows(const float* x, const float* res, long ldx, const float* __restrict__ w,
                                               const float* __restrict__ b, const float* sc, float* y, long ldy, bf16* y_hi,
                                               bf16* y_lo, float* y2, long ld2, int lane, int n_valid) {
  constexpr int G4 = D / 4, PER = (G4 + 31) / 32;
  float4 v[NR][PER];
  float scale[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    scale[r] = (sc && r < n_valid) ? sc[r] : 1.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int g = lane + 32 * i;
      v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < G4 && r < n_valid) {
        v[r][i] = *reinterpret_cast<const float4*>(x + r * ldx + g * 4);
        if (res) {
          const float4 r4 = *reinterpret_cast<const float4*>(res + r * ldx + g * 4);
          v[r][i].x += r4.x; v[r][i].y += r4.y; v[r][i].z += r4.z; v[r][i].w += r4.w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (r >= n_valid) break;  // warp-uniform
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i)
      if (lane + 32 * i < G4) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (lane + 32 * i < G4) {
        const float dx = v[r][i].x - mean, dy = v[r][i].y - mean, dz = v[r][i].z - mean, dw = v[r][i].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int g = lane + 32 * i;
      if (g < G4) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + g * 4)), b4 = __ldg(reinterpret_cast<const float4*>(b + g * 4));
        float4 o;
        o.x = ((v[r][i].x - mean) * rstd * w4.x + b4.x) * scale[r];
        o.y = ((v[r][i].y - mean) * rstd * w4.y + b4.y) * scale[r];
        o.z = ((v[r][i].z - mean) * rstd * w4.z + b4.z) * scale[r];
        o.w = ((v[r][i].w - mean) * rstd * w4.w + b4.w) * scale[r];
        *reinterpret_cast<float4*>(y + r * ldy + g * 4) = o;
        if (y2) *reinterpret_cast<float4*>(y2 + r * ld2 + g * 4) = o;
        if (y_hi) {  // split-bf16 image for a following tensor-core GEMM
          *reinterpret_cast<uint2*>(y_hi + r * ld2 + g * 4) = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
          *reinterpret_cast<uint2*>(y_lo + r * ld2 + g * 4) =
              make_uint2(pack_bf16(o.x - bf16_round(o.x), o.y - bf16_round(o.y)), pack_bf16(o.z - bf16_round(o.z), o.w - bf16_round(o.w)));
        }
      }
    }
  }
}

}  // namespace s2s
