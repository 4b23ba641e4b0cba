// Invariant Point Attention, pair part, second generation (reference src/models/net/ipa.py:177,207-215,253-257).
//
// Same math as ipa_pair_attention_kernel (ipa.cu) — pair bias linear_b(z_ij) into the logits, softmax over keys, pair
// aggregation sum_j P_ij z_ij, down_z by linearity — but organised so that the HBM stream of the pair tensor never stops:
//   * persistent CTAs (one per SM), each looping over (decoy, query) slabs; weights are staged once per CTA;
//   * a producer warp TMA-loads the next slab (L x 128 bf16 as SWIZZLE_128B blocks, plus the 8 logit rows) into one of two
//     stages while two compute groups (4 warps each, one per stage) work on the previous ones out of phase;
//   * both in-kernel GEMMs run on tcgen05 with 16-column accumulators in tensor memory:
//       bias   D1[j][hi|lo heads] = z[j][:] . Wb^T          (A = slab, K-major;  B = split-bf16 linear_b weights)
//       zsum   D2[c][hi|lo heads] = sum_j z[j][c] P[h][j]   (A = the SAME slab read MN-major; B = split-bf16 P, K-major)
//     so a slab is read from shared memory exactly twice, by the tensor pipe, and never by ld.shared;
//   * one thread per key (tcgen05.ld gives each thread its key's 8 head biases), warp-shuffle + 4-warp reductions for
//     the softmax, attention weights leave as split bf16 (P_hi for P.v, P_hi + P_lo for P.v_pts).
// Algorithmic HBM bytes per slab: L*128*2 (z) + 8*L*4 (logits in) + 2*8*L*2 (P out) + 1 KB (o_pair).
#include <cstdlib>

#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int P2_THREADS = 288;  // warps 0-3: compute group 0, warps 4-7: compute group 1, warp 8: TMA producer
constexpr int ZS_PITCH = C_Z + 4;

template <int NKB>
struct P2Layout {
  static constexpr int KP = NKB * 128;                    // padded keys
  static constexpr int Z_BYTES = NKB * 2 * TILE_BYTES;    // slab blocks [rb][cb], 16 KB each
  static constexpr int SP_BYTES = NKB * 4096;             // logits rows [8][KP] fp32, later the P operand [16][KP] bf16
  static constexpr int STAGE = Z_BYTES + SP_BYTES;
  static constexpr int OFF_WB = 2 * STAGE;                // linear_b image: 2 blocks [16 rows x 64 ch]
  static constexpr int OFF_WDZ = OFF_WB + 4096;           // down_z weight, [128][32] fp32
  static constexpr int OFF_ZS = OFF_WDZ + C_Z * 32 * 4;   // [2 groups][2 queries of a slab][8][ZS_PITCH] fp32
  static constexpr int OFF_RED = OFF_ZS + 2 * 2 * 8 * ZS_PITCH * 4;  // [2 groups][2][4 warps][8]
  static constexpr int OFF_BAR = OFF_RED + 2 * 2 * 4 * 8 * 4;
  static constexpr int BYTES = OFF_BAR + 8 * 8 + 16;
};

struct P2Args {
  const float* S;
  const float* mask;
  const float* bb;
  const float* Wdz_t;
  const float* bdz;
  const bf16* wb_img;
  bf16 *P_hi, *P_lo;
  float* o_pair;
  bf16 *opair_hi, *opair_lo;
  long ld_opair;
  int L, n_slabs;
  uint32_t a2_lbo, a2_sbo;  // MN-major descriptor fields of the zsum A operand (16-byte units)
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  tmem_wait_ld();
}

// D += A B^T with full 64-bit descriptors given as (lo, hi) words
template <bool kAccumulate>
__device__ __forceinline__ void umma_desc(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(kAccumulate ? 1 : 0)
      : "memory");
}

// QPS = queries per slab.  A slab is NKB blocks of 128 pair rows; at L = 64 one block holds the keys of TWO consecutive queries
// (their rows are adjacent in memory), and with QPS = 1 half of every block — loaded, multiplied and carried through the softmax
// by half-idle warps — belonged to the next query and was thrown away: the kernel ran at 1.2 TB/s there (one ~2 us slab step
// per 16 KB of useful data).  QPS = 2 (L = 64 only) keeps both: per-query softmax reductions over the two warps that own the
// query's keys, one zsum accumulator per query (the K range of the MMA split at the query boundary), down_z once per query.
template <int NKB, int QPS>
__global__ void __launch_bounds__(P2_THREADS, 1)
ipa_pair_tc_kernel(const __grid_constant__ CUtensorMap tmap_z, P2Args a) {
  static_assert(QPS == 1 || (QPS == 2 && NKB == 1), "two queries per slab: L = 64, one key block");
  using LY = P2Layout<NKB>;
  constexpr int KP = LY::KP;
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* wdz_s = reinterpret_cast<float*>(smem + LY::OFF_WDZ);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint64_t* full = bars;          // [2] stage loaded (TMA bytes)
  uint64_t* empty = bars + 2;     // [2] stage free (zsum MMAs retired)
  uint64_t* bias_bar = bars + 4;  // [2] bias accumulators ready
  uint64_t* zsum_bar = bars + 6;  // [2] zsum accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int L = a.L;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(&bars[s], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  for (int idx = threadIdx.x; idx < 4096 / 16; idx += P2_THREADS)
    reinterpret_cast<uint4*>(smem + LY::OFF_WB)[idx] = reinterpret_cast<const uint4*>(a.wb_img)[idx];
  for (int idx = threadIdx.x; idx < C_Z * 32 / 4; idx += P2_THREADS)
    reinterpret_cast<float4*>(wdz_s)[idx] = reinterpret_cast<const float4*>(a.Wdz_t)[idx];
  fence_proxy_async();  // the linear_b image is read by the tensor pipe (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // only weights (constant during an iteration) were read so far
  const int n_local = a.n_slabs > (int)blockIdx.x ? (a.n_slabs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int n = 0; n < n_local; ++n) {
        const int g = n & 1, k = n >> 1;
        const int slab = blockIdx.x + n * gridDim.x;
        const int q0 = slab * QPS;  // first query of the slab
        const int b = q0 / L, i = q0 - b * L;
        mbar_wait(&empty[g], (k & 1) ^ 1);
        unsigned char* st = smem + g * LY::STAGE;
        mbar_expect_tx(&full[g], LY::Z_BYTES + 8 * QPS * L * 4);
#pragma unroll
        for (int rb = 0; rb < NKB; ++rb)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
            tma_load_2d(st + (rb * 2 + cb) * TILE_BYTES, &tmap_z, cb * KBLK, q0 * L + rb * 128, &full[g]);
        const float* Srow = a.S + ((long)b * N_H * L + i) * L;
        for (int h = 0; h < N_H; ++h)  // the logits rows of consecutive queries are adjacent: one copy of QPS * L floats per head
          tma_bulk_1d(st + LY::Z_BYTES + h * KP * 4, Srow + (long)h * L * L, QPS * L * 4, &full[g]);
      }
    }
  } else {
    // ===== compute group g (bound to stage g): thread t owns keys t, t+128 (..) for the softmax and channel t for zsum =====
    const int g = warp >> 2, wq = warp & 3, t = threadIdx.x & 127;
    unsigned char* st = smem + g * LY::STAGE;
    float* S_s = reinterpret_cast<float*>(st + LY::Z_BYTES);
    unsigned char* P_s = st + LY::Z_BYTES;
    float* zs = reinterpret_cast<float*>(smem + LY::OFF_ZS) + g * 2 * 8 * ZS_PITCH;
    float* red = reinterpret_cast<float*>(smem + LY::OFF_RED) + g * 64;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t d1 = tmem + g * 64, d2 = tmem + g * 64 + 32;
    const uint32_t z_lo = desc_lo_sw128(smem_u32(st));                        // K-major view of the slab (bias MMA)
    const uint32_t z_mn = ((smem_u32(st) >> 4) & 0x3FFFu) | (a.a2_lbo << 16);  // MN-major view (zsum MMA)
    const uint32_t mn_hi = a.a2_sbo | (1u << 14) | (2u << 29);
    const uint32_t wb_lo = desc_lo_sw128(smem_u32(smem + LY::OFF_WB));
    const uint32_t p_lo = desc_lo_sw128(smem_u32(P_s));
    constexpr uint32_t IDESC_B = make_idesc(128, 16);
    constexpr uint32_t IDESC_Z = make_idesc(128, 16) | (1u << 15);  // A is MN-major
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    float bbias[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) bbias[h] = a.bb[h];
    const int dd = t & 31, hq = t >> 5;
    const float bdz = a.bdz[dd];
    constexpr float SQRT1_3 = 0.57735026918962576f;

    for (int k = 0;; ++k) {
      const int n = 2 * k + g;
      if (n >= n_local) break;
      const int slab = blockIdx.x + n * gridDim.x;
      const int q0 = slab * QPS;
      const int tq = QPS == 2 ? (t >> 6) : 0;  // which query of the slab this thread's key belongs to
      const int b = q0 / L, i = q0 - b * L + tq;
      float mk[NKB];
#pragma unroll
      for (int rb = 0; rb < NKB; ++rb) {
        const int j = QPS == 2 ? (t & 63) : rb * 128 + t;
        mk[rb] = j < L ? __ldg(a.mask + (long)b * L + j) : 0.f;
      }
      const float m_i = __ldg(a.mask + (long)b * L + i);
      mbar_wait(&full[g], k & 1);
      tc_fence_after();
      if (wq == 0) {
        if (elect_one()) {
#pragma unroll
          for (int rb = 0; rb < NKB; ++rb)
#pragma unroll
            for (int cb = 0; cb < 2; ++cb)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t al = z_lo + (rb * 2 + cb) * BLK + ks * 2, bl = wb_lo + cb * (2048 >> 4) + ks * 2;
                if (cb | ks) umma_ss<true>(d1 + rb * 16, al, bl, IDESC_B); else umma_ss<false>(d1 + rb * 16, al, bl, IDESC_B);
              }
          umma_commit(&bias_bar[g]);
        }
        __syncwarp();
      }
      float x[NKB][8];
#pragma unroll
      for (int rb = 0; rb < NKB; ++rb)
#pragma unroll
        for (int h = 0; h < 8; ++h) x[rb][h] = S_s[h * KP + rb * 128 + t];
      mbar_wait(&bias_bar[g], k & 1);
      tc_fence_after();
      float mx[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) mx[h] = -INFINITY;
#pragma unroll
      for (int rb = 0; rb < NKB; ++rb) {
        float d[16];
        tmem_ld16(d1 + lane_off + rb * 16, d);
        const bool ok = QPS == 2 || rb * 128 + t < L;
        const float mterm = 1e5f * (m_i * mk[rb] - 1.f);
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float v = ok ? x[rb][h] + SQRT1_3 * (d[h] + d[8 + h] + bbias[h]) + mterm : -INFINITY;
          x[rb][h] = v;
          mx[h] = fmaxf(mx[h], v);
        }
      }
      tc_fence_before();
#pragma unroll
      for (int h = 0; h < 8; ++h) mx[h] = warp_max(mx[h]);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) red[wq * 8 + h] = mx[h];
      }
      named_bar_sync(1 + g, 128);
      float sum[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if constexpr (QPS == 2) mx[h] = fmaxf(red[(wq & 2) * 8 + h], red[((wq & 2) + 1) * 8 + h]);  // the two warps of this query
        else mx[h] = fmaxf(fmaxf(red[h], red[8 + h]), fmaxf(red[16 + h], red[24 + h]));
        sum[h] = 0.f;
      }
#pragma unroll
      for (int rb = 0; rb < NKB; ++rb)
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float e = __expf(x[rb][h] - mx[h]);
          x[rb][h] = e;
          sum[h] += e;
        }
#pragma unroll
      for (int h = 0; h < 8; ++h) sum[h] = warp_sum(sum[h]);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) red[32 + wq * 8 + h] = sum[h];
      }
      named_bar_sync(1 + g, 128);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if constexpr (QPS == 2) sum[h] = 1.f / (red[32 + (wq & 2) * 8 + h] + red[32 + ((wq & 2) + 1) * 8 + h]);
        else sum[h] = 1.f / ((red[32 + h] + red[40 + h]) + (red[48 + h] + red[56 + h]));
      }
      // attention weights: split bf16 to HBM (A operands of P.v / P.v_pts) and into the stage as the zsum B operand
      // (rows 0-7 hi, 8-15 lo; K-major SWIZZLE_128B blocks of 64 keys), overwriting the logits rows (all read above)
#pragma unroll
      for (int rb = 0; rb < NKB; ++rb) {
        const int j = QPS == 2 ? (t & 63) : rb * 128 + t;
        unsigned char* blk = P_s + (rb * 2 + (t >> 6)) * 2048;
        const int jj = t & 63;
        bf16* gh = a.P_hi + ((long)b * N_H * L + i) * L + j;
        bf16* gl = a.P_lo + ((long)b * N_H * L + i) * L + j;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float p = x[rb][h] * sum[h];
          const bf16 hi = __float2bfloat16_rn(p);
          const bf16 lo = __float2bfloat16_rn(p - __bfloat162float(hi));
          *reinterpret_cast<bf16*>(blk + sw128_offset(h, jj)) = hi;
          *reinterpret_cast<bf16*>(blk + sw128_offset(8 + h, jj)) = lo;
          if (j < L) {
            gh[(long)h * L * L] = hi;
            gl[(long)h * L * L] = lo;
          }
        }
      }
      fence_proxy_async();
      named_bar_sync(1 + g, 128);
      if (wq == 0) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int rb = 0; rb < NKB; ++rb)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t al = z_mn + rb * 2 * BLK + ks * (2048 >> 4);
              const uint32_t bl = p_lo + (rb * 2 + (ks >> 2)) * (2048 >> 4) + (ks & 3) * 2;
              if constexpr (QPS == 2) {  // keys 0-63 (k-steps 0-3) are query 0, keys 64-127 query 1: one accumulator each
                const uint32_t dq = d2 + (ks >> 2) * 16;
                if (ks & 3) umma_desc<true>(dq, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
                else umma_desc<false>(dq, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
              } else {
                if (rb | ks) umma_desc<true>(d2, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
                else umma_desc<false>(d2, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
              }
            }
          umma_commit(&zsum_bar[g]);
          umma_commit(&empty[g]);  // slab and P operand consumed: the producer may refill this stage
        }
        __syncwarp();
      }
      mbar_wait(&zsum_bar[g], k & 1);
      tc_fence_after();
#pragma unroll
      for (int qq = 0; qq < QPS; ++qq) {
        float d[16];
        tmem_ld16(d2 + qq * 16 + lane_off, d);  // lane = channel t
#pragma unroll
        for (int h = 0; h < 8; ++h) zs[(qq * 8 + h) * ZS_PITCH + t] = d[h] + d[8 + h];
      }
      tc_fence_before();
      named_bar_sync(1 + g, 128);
      // o_pair[h][d] = down_z(sum_j P z) : sum_j P = 1, so the bias passes through (ipa.py:253-254)
#pragma unroll
      for (int qq = 0; qq < QPS; ++qq) {
        float acc0 = bdz, acc1 = bdz;
        const float* z0 = zs + (qq * 8 + hq) * ZS_PITCH;
        const float* z1 = zs + (qq * 8 + hq + 4) * ZS_PITCH;
#pragma unroll 4
        for (int c = 0; c < C_Z; c += 4) {
          const float4 u0 = *reinterpret_cast<const float4*>(z0 + c);
          const float4 u1 = *reinterpret_cast<const float4*>(z1 + c);
          const float w0 = wdz_s[c * 32 + dd], w1 = wdz_s[(c + 1) * 32 + dd], w2 = wdz_s[(c + 2) * 32 + dd], w3 = wdz_s[(c + 3) * 32 + dd];
          acc0 = fmaf(w0, u0.x, acc0); acc0 = fmaf(w1, u0.y, acc0); acc0 = fmaf(w2, u0.z, acc0); acc0 = fmaf(w3, u0.w, acc0);
          acc1 = fmaf(w0, u1.x, acc1); acc1 = fmaf(w1, u1.y, acc1); acc1 = fmaf(w2, u1.z, acc1); acc1 = fmaf(w3, u1.w, acc1);
        }
        const long o0 = ((long)q0 + qq) * a.ld_opair + hq * 32 + dd;  // row of query q0 + qq = b * L + its index
        if (a.opair_hi) {
          const bf16 h0 = __float2bfloat16_rn(acc0), h1 = __float2bfloat16_rn(acc1);
          a.opair_hi[o0] = h0;
          a.opair_hi[o0 + 128] = h1;
          a.opair_lo[o0] = __float2bfloat16_rn(acc0 - __bfloat162float(h0));
          a.opair_lo[o0 + 128] = __float2bfloat16_rn(acc1 - __bfloat162float(h1));
        } else {
          a.o_pair[o0] = acc0;
          a.o_pair[o0 + 128] = acc1;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- long chains (256 < L <= 512) -----------------------------------------------------------------------------------
// A slab no longer fits twice in shared memory, so the pipeline runs at key-block granularity: a ring of five 32 KB
// blocks ([128 keys x 128 channels]) is filled by the producer; a slab's NKB blocks stay resident from their bias MMA
// until their zsum MMA, whose per-block tcgen05.commit hands the slot straight back to the producer, which is already
// prefetching the next slab's blocks.  One 8-warp compute group (two warps per TMEM lane quarter; a thread owns keys
// rb*128 + its lane for rb = half, half + 2), a producer warp and a dedicated MMA-issuer warp; bias accumulators and
// the logits / P buffers are double-buffered by slab parity so bias(n+1) can start while slab n is in its epilogue.
constexpr int PL_THREADS = 320;
constexpr int PL_RING = 5;

template <int NKB>
struct PLLayout {
  static constexpr int KP = NKB * 128;
  static constexpr int BLOCK = 2 * TILE_BYTES;            // one key block: two 64-channel SW128 sub-blocks
  static constexpr int SP_BYTES = NKB * 4096;             // logits rows [8][KP] fp32 -> P operand [16][KP] bf16
  static constexpr int OFF_SP = PL_RING * BLOCK;
  static constexpr int OFF_WB = OFF_SP + 2 * SP_BYTES;
  static constexpr int OFF_WDZ = OFF_WB + 4096;
  static constexpr int OFF_ZS = OFF_WDZ + C_Z * 32 * 4;   // [8][ZS_PITCH]
  static constexpr int OFF_RED = OFF_ZS + 8 * ZS_PITCH * 4;  // [2][8 warps][8]
  static constexpr int OFF_BAR = OFF_RED + 2 * 8 * 8 * 4;
  static constexpr int N_BARS = 2 * PL_RING + 9;
  static constexpr int BYTES = OFF_BAR + N_BARS * 8 + 16;
};

template <int NKB>
__global__ void __launch_bounds__(PL_THREADS, 1)
ipa_pair_tc_long_kernel(const __grid_constant__ CUtensorMap tmap_z, P2Args a) {
  using LY = PLLayout<NKB>;
  constexpr int KP = LY::KP;
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* wdz_s = reinterpret_cast<float*>(smem + LY::OFF_WDZ);
  float* zs = reinterpret_cast<float*>(smem + LY::OFF_ZS);
  float* red = reinterpret_cast<float*>(smem + LY::OFF_RED);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint64_t* full = bars;                        // [PL_RING] key block loaded
  uint64_t* empty = bars + PL_RING;             // [PL_RING] key block consumed by its zsum MMAs
  uint64_t* s_full = bars + 2 * PL_RING;        // [2] logits rows loaded
  uint64_t* sp_empty = s_full + 2;              // [2] P operand consumed
  uint64_t* bias_bar = s_full + 4;              // [2] bias accumulators ready
  uint64_t* p_ready = s_full + 6;               // [2] P operand written (256 arrivals)
  uint64_t* zsum_bar = s_full + 8;              // zsum accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + LY::N_BARS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int L = a.L;
  if (threadIdx.x == 0) {
    for (int s = 0; s < LY::N_BARS; ++s) mbar_init(&bars[s], 1);
    mbar_init(&p_ready[0], 256);
    mbar_init(&p_ready[1], 256);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  for (int idx = threadIdx.x; idx < 4096 / 16; idx += PL_THREADS)
    reinterpret_cast<uint4*>(smem + LY::OFF_WB)[idx] = reinterpret_cast<const uint4*>(a.wb_img)[idx];
  for (int idx = threadIdx.x; idx < C_Z * 32 / 4; idx += PL_THREADS)
    reinterpret_cast<float4*>(wdz_s)[idx] = reinterpret_cast<const float4*>(a.Wdz_t)[idx];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // only weights (constant during an iteration) were read so far
  const int n_local = a.n_slabs > (int)blockIdx.x ? (a.n_slabs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t BLK = TILE_BYTES >> 4;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t blk = 0;
      for (int n = 0; n < n_local; ++n) {
        const int p = n & 1;
        const int slab = blockIdx.x + n * gridDim.x;
        const int b = slab / L, i = slab - b * L;
        mbar_wait(&sp_empty[p], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&s_full[p], 8 * L * 4);
        const float* Srow = a.S + ((long)b * N_H * L + i) * L;
        unsigned char* sp = smem + LY::OFF_SP + p * LY::SP_BYTES;
        for (int h = 0; h < N_H; ++h) tma_bulk_1d(sp + h * KP * 4, Srow + (long)h * L * L, L * 4, &s_full[p]);
        for (int rb = 0; rb < NKB; ++rb, ++blk) {
          const uint32_t slot = blk % PL_RING, ph = (blk / PL_RING) & 1;
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect_tx(&full[slot], LY::BLOCK);
          unsigned char* dst = smem + slot * LY::BLOCK;
          tma_load_2d(dst, &tmap_z, 0, slab * L + rb * 128, &full[slot]);
          tma_load_2d(dst + TILE_BYTES, &tmap_z, KBLK, slab * L + rb * 128, &full[slot]);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer (whole warp runs the loop, one elected lane issues) =====
    const uint32_t ring_lo = desc_lo_sw128(smem_u32(smem));
    const uint32_t ring_mn = ((smem_u32(smem) >> 4) & 0x3FFFu) | (a.a2_lbo << 16);
    const uint32_t mn_hi = a.a2_sbo | (1u << 14) | (2u << 29);
    const uint32_t wb_lo = desc_lo_sw128(smem_u32(smem + LY::OFF_WB));
    constexpr uint32_t IDESC_B = make_idesc(128, 16);
    constexpr uint32_t IDESC_Z = make_idesc(128, 16) | (1u << 15);
    uint32_t blk = 0;
    for (int n = 0; n < n_local; ++n, blk += NKB) {
      const int p = n & 1;
      const uint32_t d1 = tmem + p * 64, d2 = tmem + 128;
      for (int rb = 0; rb < NKB; ++rb) {
        const uint32_t slot = (blk + rb) % PL_RING, ph = ((blk + rb) / PL_RING) & 1;
        mbar_wait(&full[slot], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t al = ring_lo + (slot * 2 + cb) * BLK + ks * 2, bl = wb_lo + cb * (2048 >> 4) + ks * 2;
              if (cb | ks) umma_ss<true>(d1 + rb * 16, al, bl, IDESC_B); else umma_ss<false>(d1 + rb * 16, al, bl, IDESC_B);
            }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bias_bar[p]);
      __syncwarp();
      mbar_wait(&p_ready[p], (n >> 1) & 1);
      tc_fence_after();
      const uint32_t p_lo = desc_lo_sw128(smem_u32(smem + LY::OFF_SP + p * LY::SP_BYTES));
      for (int rb = 0; rb < NKB; ++rb) {
        const uint32_t slot = (blk + rb) % PL_RING;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t al = ring_mn + slot * 2 * BLK + ks * (2048 >> 4);
            const uint32_t bl = p_lo + (rb * 2 + (ks >> 2)) * (2048 >> 4) + (ks & 3) * 2;
            if (rb | ks) umma_desc<true>(d2, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
            else umma_desc<false>(d2, al, mn_hi, bl, DESC_HI_SW128, IDESC_Z);
          }
          umma_commit(&empty[slot]);  // this key block is done: the producer may refill its slot
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(zsum_bar);
        umma_commit(&sp_empty[p]);
      }
      __syncwarp();
    }
  } else {
    // ===== compute group: 8 warps, TMEM lane quarter q = warp % 4, half = warp / 4 =====
    const int q = warp & 3, half = warp >> 2, t = threadIdx.x;
    const int kl = q * 32 + lane;  // key within a block (= TMEM lane) / channel for zsum
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    constexpr int NU = (NKB + 1) / 2;  // key blocks per thread: rb = half + 2u
    float bbias[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) bbias[h] = a.bb[h];
    const int dd = t & 31, hd = t >> 5;
    const float bdz = a.bdz[dd];
    constexpr float SQRT1_3 = 0.57735026918962576f;
    for (int n = 0; n < n_local; ++n) {
      const int p = n & 1;
      const int slab = blockIdx.x + n * gridDim.x;
      const int b = slab / L, i = slab - b * L;
      float* S_s = reinterpret_cast<float*>(smem + LY::OFF_SP + p * LY::SP_BYTES);
      unsigned char* P_s = smem + LY::OFF_SP + p * LY::SP_BYTES;
      float mk[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int j = (half + 2 * u) * 128 + kl;
        mk[u] = (half + 2 * u < NKB && j < L) ? __ldg(a.mask + (long)b * L + j) : 0.f;
      }
      const float m_i = __ldg(a.mask + (long)b * L + i);
      mbar_wait(&s_full[p], (n >> 1) & 1);
      float x[NU][8];
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int h = 0; h < 8; ++h) x[u][h] = (half + 2 * u < NKB) ? S_s[h * KP + (half + 2 * u) * 128 + kl] : 0.f;
      mbar_wait(&bias_bar[p], (n >> 1) & 1);
      tc_fence_after();
      float mx[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) mx[h] = -INFINITY;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int rb = half + 2 * u;
        if (rb < NKB) {  // warp-uniform
          float d[16];
          tmem_ld16(tmem + p * 64 + rb * 16 + lane_off, d);
          const bool ok = rb * 128 + kl < L;
          const float mterm = 1e5f * (m_i * mk[u] - 1.f);
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const float v = ok ? x[u][h] + SQRT1_3 * (d[h] + d[8 + h] + bbias[h]) + mterm : -INFINITY;
            x[u][h] = v;
            mx[h] = fmaxf(mx[h], v);
          }
        } else {
#pragma unroll
          for (int h = 0; h < 8; ++h) x[u][h] = -INFINITY;
        }
      }
      tc_fence_before();
#pragma unroll
      for (int h = 0; h < 8; ++h) mx[h] = warp_max(mx[h]);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) red[warp * 8 + h] = mx[h];
      }
      named_bar_sync(1, 256);
      float sum[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        float m = red[h];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w * 8 + h]);
        mx[h] = m;
        sum[h] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < NU; ++u)
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float e = __expf(x[u][h] - mx[h]);
          x[u][h] = e;
          sum[h] += e;
        }
#pragma unroll
      for (int h = 0; h < 8; ++h) sum[h] = warp_sum(sum[h]);
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) red[64 + warp * 8 + h] = sum[h];
      }
      named_bar_sync(1, 256);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[64 + w * 8 + h];
        sum[h] = 1.f / s;
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int rb = half + 2 * u;
        if (rb >= NKB) continue;
        const int j = rb * 128 + kl;
        unsigned char* blkp = P_s + (rb * 2 + (kl >> 6)) * 2048;
        const int jj = kl & 63;
        bf16* gh = a.P_hi + ((long)b * N_H * L + i) * L + j;
        bf16* gl = a.P_lo + ((long)b * N_H * L + i) * L + j;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float pv = x[u][h] * sum[h];
          const bf16 hi = __float2bfloat16_rn(pv);
          const bf16 lo = __float2bfloat16_rn(pv - __bfloat162float(hi));
          *reinterpret_cast<bf16*>(blkp + sw128_offset(h, jj)) = hi;
          *reinterpret_cast<bf16*>(blkp + sw128_offset(8 + h, jj)) = lo;
          if (j < L) {
            gh[(long)h * L * L] = hi;
            gl[(long)h * L * L] = lo;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&p_ready[p]);
      mbar_wait(zsum_bar, n & 1);
      tc_fence_after();
      if (half == 0) {
        float d[16];
        tmem_ld16(tmem + 128 + lane_off, d);  // lane = channel kl
#pragma unroll
        for (int h = 0; h < 8; ++h) zs[h * ZS_PITCH + kl] = d[h] + d[8 + h];
      }
      tc_fence_before();
      named_bar_sync(1, 256);
      float acc = bdz;
      const float* z0 = zs + hd * ZS_PITCH;
#pragma unroll 8
      for (int c = 0; c < C_Z; c += 4) {
        const float4 u0 = *reinterpret_cast<const float4*>(z0 + c);
        acc = fmaf(wdz_s[c * 32 + dd], u0.x, acc);
        acc = fmaf(wdz_s[(c + 1) * 32 + dd], u0.y, acc);
        acc = fmaf(wdz_s[(c + 2) * 32 + dd], u0.z, acc);
        acc = fmaf(wdz_s[(c + 3) * 32 + dd], u0.w, acc);
      }
      const long o0 = ((long)b * L + i) * a.ld_opair + hd * 32 + dd;
      if (a.opair_hi) {
        const bf16 h0 = __float2bfloat16_rn(acc);
        a.opair_hi[o0] = h0;
        a.opair_lo[o0] = __float2bfloat16_rn(acc - __bfloat162float(h0));
      } else {
        a.o_pair[o0] = acc;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// [Wb_hi (8 rows); Wb_lo (8 rows)] x 128 channels -> two [16 x 64] bf16 blocks in the SW128 K-major layout
__global__ void build_ipa_wb_kernel(const float* __restrict__ Wb, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 16 * C_Z) return;
  const int r = idx / C_Z, c = idx % C_Z;
  const float w = Wb[(r & 7) * C_Z + c];
  const bf16 hi = __float2bfloat16_rn(w);
  const bf16 v = r < 8 ? hi : __float2bfloat16_rn(w - __bfloat162float(hi));
  *reinterpret_cast<bf16*>(dst + (c / KBLK) * 2048 + sw128_offset(r, c % KBLK)) = v;
}

template <int NKB, int QPS>
void launch_pair_tc(const IpaPairArgs& a, const CUtensorMap& mz, const P2Args& k, cudaStream_t st) {
  using LY = P2Layout<NKB>;
  static bool configured = false;
  const int smem = LY::BYTES + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(ipa_pair_tc_kernel<NKB, QPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int cap = sm_count() * (NKB == 1 ? 2 : 1);
  launch_pdl(ipa_pair_tc_kernel<NKB, QPS>, k.n_slabs < cap ? k.n_slabs : cap, P2_THREADS, smem, st, mz, k);
  S2S_LAUNCH_CHECK();
}

template <int NKB>
void launch_pair_tc_long(const CUtensorMap& mz, const P2Args& k, cudaStream_t st) {
  using LY = PLLayout<NKB>;
  static bool configured = false;
  const int smem = LY::BYTES;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(ipa_pair_tc_long_kernel<NKB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  launch_pdl(ipa_pair_tc_long_kernel<NKB>, k.n_slabs < sm_count() ? k.n_slabs : sm_count(), PL_THREADS, smem, st, mz, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace

size_t ipa_wb_img_elems() { return 2 * 16 * KBLK; }

void build_ipa_wb_img(const float* Wb, bf16* dst, cudaStream_t st) {
  build_ipa_wb_kernel<<<ceil_div(16 * C_Z, 256), 256, 0, st>>>(Wb, reinterpret_cast<unsigned char*>(dst));
  S2S_LAUNCH_CHECK();
}

bool ipa_pair_attention_tc_supported(int L) { return L >= 1 && L <= 512 && L % 4 == 0; }

void ipa_pair_attention_tc(const IpaPairArgs& a, cudaStream_t st) {
  S2S_CHECK(ipa_pair_attention_tc_supported(a.L), "ipa_pair_attention_tc: needs L <= 512 and L % 4 == 0");
  S2S_CHECK(a.wb_img && a.P_bf16 && a.P_lo, "ipa_pair_attention_tc: weight image / P outputs missing");
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z, rows, C_Z, C_Z);
  P2Args k;
  k.S = a.S; k.mask = a.mask; k.bb = a.bb; k.Wdz_t = a.Wdz_t; k.bdz = a.bdz; k.wb_img = a.wb_img;
  k.P_hi = a.P_bf16; k.P_lo = a.P_lo; k.o_pair = a.o_pair; k.ld_opair = a.ld_opair;
  k.opair_hi = a.opair_hi; k.opair_lo = a.opair_lo;
  k.L = a.L; k.n_slabs = a.B * a.L;
  k.a2_lbo = TILE_BYTES >> 4;  // between the two 64-channel blocks of a key block
  k.a2_sbo = 1024 >> 4;        // between 8-key groups
  if (const char* e = getenv("S2S_IPA_DEBUG")) {  // descriptor experiments only
    if (atoi(e) & 1) { const uint32_t tmp = k.a2_lbo; k.a2_lbo = k.a2_sbo; k.a2_sbo = tmp; }
  }
  S2S_PROF("ipa_pair_attention", st);
  static const int qps_env = [] { const char* e = getenv("S2S_IPA_QPS"); return e ? atoi(e) : 2; }();  // 1: A/B timing
  if (a.L == 64 && qps_env == 2) {  // two queries per slab (see the kernel)
    k.n_slabs = a.B * a.L / 2;
    launch_pair_tc<1, 2>(a, mz, k, st);
  } else if (a.L <= 128) launch_pair_tc<1, 1>(a, mz, k, st);
  else if (a.L <= 256) launch_pair_tc<2, 1>(a, mz, k, st);
  else if (a.L <= 384) launch_pair_tc_long<3>(mz, k, st);
  else launch_pair_tc_long<4>(mz, k, st);
}

}  // namespace s2s
