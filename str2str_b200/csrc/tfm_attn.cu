// Sequence-transformer self-attention (nn.MultiheadAttention inside the reference's TransformerEncoder, ipa.py:312-317,357:
// d_model 320, 4 heads of 80, additive float key-padding mask), fused: one CTA computes softmax(q k^T / sqrt(80) + keybias) v
// for 128 queries of one (decoy, head) without the logits or the attention weights ever leaving the SM.
//
//   TMA   q [128 x 80], k [L x 80], v [L x 80] straight out of the in_proj output (bf16, row-major [B*L][960]) as SW128 boxes;
//   MMA 1 S[128 x 256] = q k^T          (tcgen05, M128 N256 K80: five K-steps; accumulator = 256 TMEM columns)
//   softmax: 8 warps, one row per thread and half of the keys each; two passes over the accumulator (max, then exp / sum);
//         the un-normalised weights go back INTO the accumulator's columns as packed bf16 (each 32-column chunk is
//         overwritten by the thread that just read it), which makes them the TS-mode A operand of
//   MMA 2 O[128 x 80] = P v             (A from tensor memory, B = v read MN-major: no transposed copy of v exists)
//   epilogue: O / rowsum -> fp32 rows + split-bf16 image for the out_proj GEMM.
// Single-pass bf16 operands: measured on full trajectories (tools/traj_parity.py), the transformer's attention does not
// need the split-bf16 treatment of the residual-stream GEMMs.  Requires L <= 256 and L % 16 == 0; other shapes use the
// GEMM + softmax + GEMM path.
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int TA_THREADS = 288;                         // warps 0-7: softmax / epilogue, warp 8: TMA + MMA issue
constexpr int TA_OFF_Q = 0;                             // 2 boxes [128 x 64]
constexpr int TA_OFF_K = 2 * TILE_BYTES;                // 2 column boxes x [256 x 64]
constexpr int TA_OFF_V = TA_OFF_K + 4 * TILE_BYTES;     // 4 key blocks x 2 column boxes x [64 x 64]
constexpr int TA_OFF_KB = TA_OFF_V + 4 * TILE_BYTES;    // key bias, 256 floats
constexpr int TA_OFF_RED = TA_OFF_KB + 256 * 4;         // [2 stats][2 halves][128 rows]
constexpr int TA_OFF_BAR = TA_OFF_RED + 4 * 128 * 4;
constexpr int TA_SMEM = TA_OFF_BAR + 4 * 8 + 16;
constexpr uint32_t TA_COL_O = 256;

struct TaArgs {
  const float* keybias;  // [B*L]
  float* y;              // [B*L][320]
  bf16 *y_hi, *y_lo;
  int L, MT;
  float scale;
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// D += A[tmem] B[smem]^T with an explicit B descriptor (MN-major B)
template <bool kAccumulate>
__device__ __forceinline__ void umma_ts_desc(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(kAccumulate ? 1 : 0)
      : "memory");
}

__global__ void __launch_bounds__(TA_THREADS, 1)
tfm_attention_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v, TaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* kb_s = reinterpret_cast<float*>(smem + TA_OFF_KB);
  float* red_s = reinterpret_cast<float*>(smem + TA_OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TA_OFF_BAR);
  uint64_t* full = bars;       // operands loaded
  uint64_t* s_full = bars + 1; // logits accumulator ready
  uint64_t* p_ready = bars + 2;  // attention weights written (256 arrivals)
  uint64_t* o_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int L = a.L;
  const int mt = blockIdx.x % a.MT, h = (blockIdx.x / a.MT) % TFM_H, b = blockIdx.x / (a.MT * TFM_H);
  const int row0 = b * L;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  pdl_sync();
  if (threadIdx.x < 256) kb_s[threadIdx.x] = threadIdx.x < L ? a.keybias[row0 + threadIdx.x] * 1.4426950408889634f : 0.f;  // log2 e folded in
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int NKB = (L + 63) / 64;  // 64-key blocks that hold real keys

  if (warp == 8) {
    if (elect_one()) {
      mbar_expect_tx(full, (2 + 4) * TILE_BYTES + NKB * TILE_BYTES);
      for (int cb = 0; cb < 2; ++cb) {
        tma_load_2d(smem + TA_OFF_Q + cb * TILE_BYTES, &map_qk, h * TFM_HD + cb * 64, row0 + mt * 128, full);
        for (int rb = 0; rb < 2; ++rb)
          tma_load_2d(smem + TA_OFF_K + (cb * 2 + rb) * TILE_BYTES, &map_qk, D_TFM + h * TFM_HD + cb * 64, row0 + rb * 128, full);
      }
      for (int kbk = 0; kbk < NKB; ++kbk)
        for (int cb = 0; cb < 2; ++cb)
          tma_load_2d(smem + TA_OFF_V + kbk * TILE_BYTES + cb * (TILE_BYTES / 2), &map_v, 2 * D_TFM + h * TFM_HD + cb * 64, row0 + kbk * 64, full);
    }
    __syncwarp();
    mbar_wait(full, 0);
    tc_fence_after();
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    const uint32_t q_lo = desc_lo_sw128(smem_u32(smem + TA_OFF_Q)), k_lo = desc_lo_sw128(smem_u32(smem + TA_OFF_K));
    if (elect_one()) {  // S = q k^T: K-steps 0-3 from column box 0, K-step 4 (columns 64..79) from box 1
      constexpr uint32_t IDESC_S = make_idesc(128, 256);
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        const uint32_t al = q_lo + (ks >> 2) * BLK + (ks & 3) * 2, bl = k_lo + (ks >> 2) * 2 * BLK + (ks & 3) * 2;
        if (ks) umma_ss<true>(tmem, al, bl, IDESC_S); else umma_ss<false>(tmem, al, bl, IDESC_S);
      }
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_ready, 0);
    tc_fence_after();
    if (elect_one()) {  // O = P v: A = packed weights in the accumulator's own columns, B = v (MN-major, N = 80)
      constexpr uint32_t IDESC_O = make_idesc(128, TFM_HD) | (1u << 16);
      const uint32_t v_lo = ((smem_u32(smem + TA_OFF_V) >> 4) & 0x3FFFu) | ((uint32_t)(TILE_BYTES / 2 >> 4) << 16);
      const int nks = (L + 15) / 16;
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t acol = tmem + (ks >> 1) * 32 + (ks & 1) * 8;
        const uint32_t bl = v_lo + (ks >> 2) * BLK + (ks & 3) * (2048 >> 4);
        if (ks) umma_ts_desc<true>(tmem + TA_COL_O, acol, bl, DESC_HI_SW128, IDESC_O);
        else umma_ts_desc<false>(tmem + TA_COL_O, acol, bl, DESC_HI_SW128, IDESC_O);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3, half = warp >> 2, r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_col = tmem + lane_off + half * 128;
    mbar_wait(s_full, 0);
    tc_fence_after();
    // Everything in the exp2 domain: logit2 = s * (scale * log2 e) + keybias * log2 e (the bias is pre-multiplied in shared
    // memory), so an element costs one FFMA + FMNMX in the first pass and FADD + FFMA + EX2 + FADD in the second.  L is a
    // multiple of 16: a 32-key chunk is full, half full (first 16 keys) or absent — no per-element bounds checks.
    float v[32];
    const float s2 = a.scale * 1.4426950408889634f;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int key0 = half * 128 + c * 32;
      if (key0 >= L) break;  // warp-uniform
      const bool whole = key0 + 32 <= L;
      tmem_ld32(s_col + c * 32, v);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        if (e < 16 || whole) {
          const float4 kb = *reinterpret_cast<const float4*>(kb_s + key0 + e);
          mx = fmaxf(mx, fmaf(v[e], s2, kb.x));
          mx = fmaxf(mx, fmaf(v[e + 1], s2, kb.y));
          mx = fmaxf(mx, fmaf(v[e + 2], s2, kb.z));
          mx = fmaxf(mx, fmaf(v[e + 3], s2, kb.w));
        }
      }
    }
    red_s[half * 128 + r] = mx;
    named_bar_sync(2 + q, 64);
    mx = fmaxf(mx, red_s[(half ^ 1) * 128 + r]);
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int key0 = half * 128 + c * 32;
      if (key0 >= L) break;
      const bool whole = key0 + 32 <= L;
      tmem_ld32(s_col + c * 32, v);
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        if (e < 16 || whole) {
          const float4 kb = *reinterpret_cast<const float4*>(kb_s + key0 + e);
          const float e0 = ex2_approx(fmaf(v[e], s2, kb.x - mx)), e1 = ex2_approx(fmaf(v[e + 1], s2, kb.y - mx));
          const float e2 = ex2_approx(fmaf(v[e + 2], s2, kb.z - mx)), e3 = ex2_approx(fmaf(v[e + 3], s2, kb.w - mx));
          sum += (e0 + e1) + (e2 + e3);
          pk[e >> 1] = pack_bf16(e0, e1);
          pk[(e >> 1) + 1] = pack_bf16(e2, e3);
        } else {
          pk[e >> 1] = 0u;
          pk[(e >> 1) + 1] = 0u;
        }
      }
      tmem_st16(s_col + c * 32, pk);  // in place: the first 16 columns of the chunk this thread has just consumed
    }
    red_s[256 + half * 128 + r] = sum;
    tc_fence_before();
    mbar_arrive(p_ready);
    named_bar_sync(2 + q, 64);
    const float inv = 1.f / (sum + red_s[256 + (half ^ 1) * 128 + r]);
    mbar_wait(o_full, 0);
    tc_fence_after();
    float o[40];
    tmem_ld32_issue(tmem + lane_off + TA_COL_O + half * 40, o);
    tmem_ld8(tmem + lane_off + TA_COL_O + half * 40 + 32, o + 32);
    tmem_wait_ld();
    const int qi = mt * 128 + r;
    if (qi < L) {
      const long off = (long)(row0 + qi) * D_TFM + h * TFM_HD + half * 40;
#pragma unroll
      for (int e = 0; e < 40; e += 4) {
        const float x0 = o[e] * inv, x1 = o[e + 1] * inv, x2 = o[e + 2] * inv, x3 = o[e + 3] * inv;
        *reinterpret_cast<float4*>(a.y + off + e) = make_float4(x0, x1, x2, x3);
        *reinterpret_cast<uint2*>(a.y_hi + off + e) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
        *reinterpret_cast<uint2*>(a.y_lo + off + e) =
            make_uint2(pack_bf16(x0 - bf16_round(x0), x1 - bf16_round(x1)), pack_bf16(x2 - bf16_round(x2), x3 - bf16_round(x3)));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

bool tfm_attention_supported(int L) { return L >= 16 && L <= 256 && L % 16 == 0; }

// qkv: in_proj output [B*L][960] bf16 (q | k | v, heads of 80 inside each third); y [B*L][320] fp32 + split-bf16 image
void tfm_attention(const bf16* qkv, const float* keybias, float* y, bf16* y_hi, bf16* y_lo, int B, int L, float scale, cudaStream_t st) {
  S2S_CHECK(tfm_attention_supported(L), "tfm_attention: needs L <= 256 and L % 16 == 0");
  const size_t R = (size_t)B * L;
  const CUtensorMap mqk = make_bf16_2d_map(qkv, R, 3 * D_TFM, 3 * D_TFM, 128);
  const CUtensorMap mv = make_bf16_2d_map(qkv, R, 3 * D_TFM, 3 * D_TFM, 64);
  TaArgs k;
  k.keybias = keybias; k.y = y; k.y_hi = y_hi; k.y_lo = y_lo; k.L = L; k.MT = ceil_div(L, 128); k.scale = scale;
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(tfm_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM));
    configured = true;
  }
  S2S_PROF("tfm_attention", st);
  launch_pdl(tfm_attention_kernel, B * TFM_H * k.MT, TA_THREADS, TA_SMEM, st, mqk, mv, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
