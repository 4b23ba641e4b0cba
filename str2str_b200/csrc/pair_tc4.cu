// Fused edge embedder (reference src/models/net/denoising_ipa.py:126-158, geo_utils.py:44-56).
//
// A kernel that walks every 128-row tile through its five steps in lock-step (table gather -> MMA -> bias/ReLU restage ->
// MMA -> LayerNorm/store) exposes each step's latency (the round-1 first cut: ~17 % of either roofline, 1.16 ms at B = 64,
// L = 256: 1.07 GB written, 0.27 TFLOP).  Here the steps are stations of a pipeline, each owned by its own warps and each
// working on a DIFFERENT tile at any moment:
//
//   G  (7 warps)  layer 1 by table lookups for tile t+2  -> A1[stage]  (bf16, K-major SWIZZLE_128B, 2 stages)
//   M  (1 warp)   tcgen05.mma  layer 2 of tile t+1 (A1 x W2 -> acc2[stage]), layer 3 of tile t (A2 x W3 -> acc3[stage])
//   E2 (4 warps)  acc2 -> + b2, ReLU -> A2[stage]                                (2 stages)
//   E3 (4 warps)  acc3 -> + b3, LayerNorm (two-pass, re-reading tensor memory), mask -> z (bf16) in HBM
//
// One persistent CTA per SM (tiles strided by the grid, so the SMs write neighbouring tiles at any moment); the four
// accumulators fill the 512 columns of tensor memory; W2 / W3 stay resident in shared memory; stations hand tiles over
// through mbarriers only.  Same arithmetic and rounding points as pair_simt.cu, except that the LayerNorm statistics are
// taken in one shifted pass (sum and sum of squares of y - y_0).
//
// Tiles are 128 consecutive rows of the flattened pair tensor [B*L*L][128].  FLAT = false (L % 128 == 0): a tile lies in one
// (decoy, i) row.  FLAT = true (any other L with B*L*L % 128 == 0; api.cu pads chains to a multiple of 32): every row
// carries its own (decoy, i, j), so short chains fill their tiles with several i rows.
//
// What bounded the first cut of this pipeline (ncu, profiles/r01c_gemm_and_embedder_experiments.log): the LSU data pipe at 76 % of its
// wavefront rate, not latency.  Row-per-thread 16-byte global stores cost 32 wavefronts each (32 different 128-byte
// lines), and every warp-uniform shared-memory read of a bias / LayerNorm parameter costs 2.  So the output rows are
// turned through a swizzled shared-memory buffer and leave as 512-byte contiguous pieces, and b2 / b3 / ln_w / ln_b sit
// in the constant bank where they are instruction operands rather than loads.
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int P_G_WARPS = 7;  // 16 warps in all: 512 threads keep 128 registers per thread (544 threads are capped at 96 and spill)
constexpr int P_THREADS = 32 * (1 + P_G_WARPS + 4 + 4);
constexpr int P_WD_PITCH = C_Z + 4;
constexpr int P_OFF_A1 = 0;                    // 2 stages x 2 K-blocks
constexpr int P_OFF_A2 = 4 * TILE_BYTES;       // 1 stage x 2 K-blocks (E2 idles two thirds of the time: no second stage needed)
constexpr int P_OFF_OUT = 6 * TILE_BYTES;      // E3 output staging: 4 warps x [32 rows x 256 B], 16-byte chunks swizzled by row
constexpr int P_OFF_W = 8 * TILE_BYTES;        // W2 k0, W2 k1, W3 k0, W3 k1
constexpr int P_OFF_VEC = 12 * TILE_BYTES;
constexpr int P_VEC_FLOATS = N_BINS * P_WD_PITCH + 32;  // Wd, bin edges
constexpr int P_OFF_BAR = P_OFF_VEC + P_VEC_FLOATS * 4;
constexpr int P_SMEM = P_OFF_BAR + 20 * 8 + 16;

__constant__ float c_ee[4][C_Z];  // b2, b3, ln_w, ln_b (copied from EdgeEmbedArgs::vec4 in stream order before every launch)

// Waits of the stations that idle by design (E2, E3, M) back off instead of spinning: a spinning warp takes issue slots
// from the G warps that share its scheduler (ncu: ~2500 try_wait iterations per tile, `not selected` stalls on G).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* b, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(40);
  }
}

// pair_distogram_bin (s2s_internal.cuh) without the 22-step scan: the edges are an arithmetic progression, so a guess from
// the quotient is off by at most one either way; the two comparisons that settle it use the table's own fp32 edges, so the
// strict inequalities of geo_utils.py:44-56 are decided exactly as in the scan.
__device__ __forceinline__ int pair_distogram_bin_fast(const float* a, const float* b, const float* lower, float inv_step) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  int k = (int)fminf(fmaxf((d - lower[0]) * inv_step, 0.f), (float)(N_BINS - 1));
  if (!(d > lower[k])) --k;
  else if (k < N_BINS - 1 && d > lower[k + 1]) ++k;
  if (k < 0) return -1;
  const float upper = (k < N_BINS - 1) ? lower[k + 1] : 1e8f;
  return d < upper ? k : -1;
}

struct EePipeArgs {
  EdgeEmbedArgs e;
  const bf16* wimg;
  int n_tiles;
};

template <bool FLAT>
__global__ void __launch_bounds__(P_THREADS, 1) edge_embed_pipe_kernel(EePipeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* wd_s = reinterpret_cast<float*>(smem + P_OFF_VEC);
  float* edge_s = wd_s + N_BINS * P_WD_PITCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* a1_full = bars + 1;     // [2] G -> M
  uint64_t* a1_empty = bars + 3;    // [2] M -> G   (tcgen05.commit)
  uint64_t* acc2_full = bars + 5;   // [2] M -> E2  (tcgen05.commit)
  uint64_t* acc2_empty = bars + 7;  // [2] E2 -> M
  uint64_t* a2_full = bars + 9;     // [2] E2 -> M
  uint64_t* a2_empty = bars + 11;   // [2] M -> E2  (tcgen05.commit)
  uint64_t* acc3_full = bars + 13;  // [2] M -> E3  (tcgen05.commit)
  uint64_t* acc3_empty = bars + 15; // [2] E3 -> M
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  const EdgeEmbedArgs& e = a.e;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a1_full[s], 32 * P_G_WARPS);
      mbar_init(&a1_empty[s], 1);
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], 128);
      mbar_init(&acc3_full[s], 1);
      mbar_init(&acc3_empty[s], 128);
    }
    mbar_init(&a2_full[0], 128);
    mbar_init(&a2_empty[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int c = threadIdx.x; c < N_BINS * C_Z; c += blockDim.x) wd_s[(c / C_Z) * P_WD_PITCH + (c % C_Z)] = e.Wd[c];
  if (threadIdx.x < N_BINS) edge_s[threadIdx.x] = e.bin_lower[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_i = FLAT ? 1 : e.L / TM;
  const int n_local = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr uint32_t IDESC = make_idesc(128, 128);

  if (warp == 0) {
    // ---- M: weight load, then the MMA issue loop (whole warp converged, one elected lane issues) ----
    if (elect_one()) {
      mbar_expect_tx(w_full, 4 * TILE_BYTES);
      for (int t = 0; t < 4; ++t) tma_bulk_1d(smem + P_OFF_W + t * TILE_BYTES, a.wimg + (size_t)t * (TILE_BYTES / 2), TILE_BYTES, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    const uint32_t a1 = desc_lo_sw128(smem_u32(smem + P_OFF_A1)), a2 = desc_lo_sw128(smem_u32(smem + P_OFF_A2));
    const uint32_t wb = desc_lo_sw128(smem_u32(smem + P_OFF_W));
    // skewed by one tile: layer 2 of tile k is issued before layer 3 of tile k-1, whose operand E2 is still producing
    for (int k = 0; k <= n_local; ++k) {
      if (k < n_local) {
        const uint32_t s = k & 1, ph = (k >> 1) & 1;
        mbar_wait_backoff(&a1_full[s], ph);
        mbar_wait_backoff(&acc2_empty[s], ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          kblock_ss(tmem + s * 128, a1 + s * 2 * BLK, wb, IDESC, true);
          kblock_ss(tmem + s * 128, a1 + s * 2 * BLK + BLK, wb + BLK, IDESC, false);
          umma_commit(&a1_empty[s]);
          umma_commit(&acc2_full[s]);
        }
        __syncwarp();
      }
      if (k >= 1) {
        const int j = k - 1;
        const uint32_t s = j & 1, ph = (j >> 1) & 1;
        mbar_wait_backoff(&a2_full[0], j & 1);
        mbar_wait_backoff(&acc3_empty[s], ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          kblock_ss(tmem + 256 + s * 128, a2, wb + 2 * BLK, IDESC, true);
          kblock_ss(tmem + 256 + s * 128, a2 + BLK, wb + 3 * BLK, IDESC, false);
          umma_commit(&a2_empty[0]);
          umma_commit(&acc3_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp <= P_G_WARPS) {
    // ---- G: layer 1 by table lookups.  Warp g owns 18 consecutive rows of the tile (the last warp 20); lanes first classify those rows
    // (distogram bin, relative-position offset), then the warp walks the rows together so that every table row is one
    // coalesced 512-byte read (lane = 4 channels).
    const int g = warp - 1;
    const int c = lane * 4;
    unsigned char* const dst0 = smem + P_OFF_A1 + (c / KBLK) * TILE_BYTES;
    const float inv_step = (float)(N_BINS - 1) / (edge_s[N_BINS - 1] - edge_s[0]);
    for (int k = 0; k < n_local; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int r_begin = g * 18, r_cnt = g == P_G_WARPS - 1 ? TM - r_begin : 18;
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      unsigned char* const dst = dst0 + s * 2 * TILE_BYTES;
      if constexpr (FLAT) {
        // lane l classifies flattened row tile*128 + r_begin + l: its (decoy, i) row bi, its key row bj, bin and offset
        const long f = (long)tile * TM + r_begin + (lane < r_cnt ? lane : 0);
        const int bi_l = (int)(f / e.L);
        const int bj_l = (bi_l / e.L) * e.L + (int)(f - (long)bi_l * e.L);
        const int bin_l = pair_distogram_bin_fast(e.sc_ca + (size_t)bi_l * 3, e.sc_ca + (size_t)bj_l * 3, edge_s, inv_step);
        const int off_l = min(max((int)(e.ridx[bi_l] - e.ridx[bj_l]) - e.d_min, 0), e.n_off - 1);
        int bi_cur = __shfl_sync(0xffffffffu, bi_l, 0);
        float4 tiv = __ldg(reinterpret_cast<const float4*>(e.Ti + (size_t)bi_cur * C_Z + c));
        mbar_wait(&a1_empty[s], ph ^ 1);
#pragma unroll 6
        for (int r16 = 0; r16 < r_cnt; ++r16) {
          const int bin = __shfl_sync(0xffffffffu, bin_l, r16);
          const int off = __shfl_sync(0xffffffffu, off_l, r16);
          const int bi = __shfl_sync(0xffffffffu, bi_l, r16);
          const int bj = __shfl_sync(0xffffffffu, bj_l, r16);
          if (bi != bi_cur) {  // warp-uniform: the tile moved on to the next i row
            bi_cur = bi;
            tiv = __ldg(reinterpret_cast<const float4*>(e.Ti + (size_t)bi_cur * C_Z + c));
          }
          const float4 tjv = __ldg(reinterpret_cast<const float4*>(e.Tj + (size_t)bj * C_Z + c));
          const float4 tpv = __ldg(reinterpret_cast<const float4*>(e.Tpos + (size_t)off * C_Z + c));
          float4 h = make_float4(tiv.x + tjv.x + tpv.x, tiv.y + tjv.y + tpv.y, tiv.z + tjv.z + tpv.z, tiv.w + tjv.w + tpv.w);
          if (bin >= 0) {
            const float4 wv = *reinterpret_cast<const float4*>(wd_s + bin * P_WD_PITCH + c);
            h.x += wv.x; h.y += wv.y; h.z += wv.z; h.w += wv.w;
          }
          *reinterpret_cast<uint2*>(dst + sw128_offset(r_begin + r16, c % KBLK)) =
              make_uint2(pack_bf16(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f)), pack_bf16(fmaxf(h.z, 0.f), fmaxf(h.w, 0.f)));
        }
      } else {
        const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
        const int b = bi / e.L;
        const size_t bj0 = (size_t)b * e.L + j0 + r_begin;
        const size_t bjl = bj0 + (lane < r_cnt ? lane : 0);
        const int bin_l = pair_distogram_bin_fast(e.sc_ca + (size_t)bi * 3, e.sc_ca + bjl * 3, edge_s, inv_step);
        const int off_l = min(max((int)(e.ridx[bi] - e.ridx[bjl]) - e.d_min, 0), e.n_off - 1);
        const float4 tiv = __ldg(reinterpret_cast<const float4*>(e.Ti + (size_t)bi * C_Z + c));
        mbar_wait(&a1_empty[s], ph ^ 1);
#pragma unroll 6
        for (int r16 = 0; r16 < r_cnt; ++r16) {
          const int bin = __shfl_sync(0xffffffffu, bin_l, r16);
          const int off = __shfl_sync(0xffffffffu, off_l, r16);
          const float4 tjv = __ldg(reinterpret_cast<const float4*>(e.Tj + (bj0 + r16) * C_Z + c));
          const float4 tpv = __ldg(reinterpret_cast<const float4*>(e.Tpos + (size_t)off * C_Z + c));
          float4 h = make_float4(tiv.x + tjv.x + tpv.x, tiv.y + tjv.y + tpv.y, tiv.z + tjv.z + tpv.z, tiv.w + tjv.w + tpv.w);
          if (bin >= 0) {
            const float4 wv = *reinterpret_cast<const float4*>(wd_s + bin * P_WD_PITCH + c);
            h.x += wv.x; h.y += wv.y; h.z += wv.z; h.w += wv.w;
          }
          *reinterpret_cast<uint2*>(dst + sw128_offset(r_begin + r16, c % KBLK)) =
              make_uint2(pack_bf16(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f)), pack_bf16(fmaxf(h.z, 0.f), fmaxf(h.w, 0.f)));
        }
      }
      fence_proxy_async();
      mbar_arrive(&a1_full[s]);
    }
  } else if (warp <= P_G_WARPS + 4) {
    // ---- E2: acc2 -> + b2, ReLU -> A2 (this thread: row r of the tile, all 128 columns in two halves) ----
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    unsigned char* const abuf = smem + P_OFF_A2;
    for (int k = 0; k < n_local; ++k) {
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      mbar_wait_backoff(&acc2_full[s], ph);
      tc_fence_after();
      float v[64];
      tmem_ld32_issue(lane_base + s * 128, v);
      tmem_ld32_issue(lane_base + s * 128 + 32, v + 32);
      tmem_wait_ld();
      mbar_wait_backoff(&a2_empty[0], (k & 1) ^ 1);  // layer 3 of the previous tile has read A2
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 1) {
          tmem_ld32_issue(lane_base + s * 128 + 64, v);
          tmem_ld32_issue(lane_base + s * 128 + 96, v + 32);
          tmem_wait_ld();
          tc_fence_before();  // all of acc2[s] is in registers: it can take the tile after next
          mbar_arrive(&acc2_empty[s]);
        }
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          float h[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) h[u] = fmaxf(v[gq * 8 + u] + c_ee[0][hf * 64 + gq * 8 + u], 0.f);
          store8_sw128(abuf + hf * TILE_BYTES, r, gq * 8, h);
        }
      }
      fence_proxy_async();
      mbar_arrive(&a2_full[0]);
    }
  } else {
    // ---- E3: acc3 -> + b3, LayerNorm, mask -> z.  Statistics in one pass over tensor memory (shifted by the row's first
    // value: sum and sum of squares of y - y_0), the second pass normalises into the staging buffer, then the warp's
    // 32 rows x 256 B leave as sixteen 512-byte contiguous stores.
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_base = tmem + 256 + ((uint32_t)(q * 32) << 16);
    unsigned char* const obuf = smem + P_OFF_OUT + q * (32 * 256);
    for (int k = 0; k < n_local; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      int bi, jr;
      if constexpr (FLAT) {
        const long f = (long)tile * TM + r;
        bi = (int)(f / e.L);
        jr = (int)(f - (long)bi * e.L);
      } else {
        bi = tile / tiles_per_i;
        jr = (tile % tiles_per_i) * TM + r;
      }
      const size_t bj = (size_t)(bi / e.L) * e.L + jr;
      const float m = __ldg(e.mask + bi) * __ldg(e.mask + bj);
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      mbar_wait_backoff(&acc3_full[s], ph);
      tc_fence_after();
      const uint32_t acc = lane_base + s * 128;
      float v[32];
      float sum = 0.f, sq = 0.f, shift = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc + c0, v);
        if (c0 == 0) shift = v[0] + c_ee[1][0];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float d = v[u] + c_ee[1][c0 + u] - shift;
          sum += d;
          sq = fmaf(d, d, sq);
        }
      }
      const float dm = sum * (1.f / C_Z);                   // mean - shift
      const float rstd = rsqrtf(fmaxf(sq * (1.f / C_Z) - dm * dm, 0.f) + 1e-5f);
      const float off = shift + dm;                        // mean
#pragma unroll
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc + c0, v);
        if (c0 == C_Z - 32) {  // last read of this accumulator
          tc_fence_before();
          mbar_arrive(&acc3_empty[s]);
        }
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float o[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int cc = c0 + gq * 8 + u;
            o[u] = fmaf((v[gq * 8 + u] + c_ee[1][cc] - off) * c_ee[2][cc], rstd, c_ee[3][cc]) * m;  // no per-tile-invariant subexpression to hoist (128 registers)
          }
          const int chunk = (c0 >> 3) + gq;                // 16-byte chunk 0..15 of this row
          *reinterpret_cast<uint4*>(obuf + lane * 256 + ((chunk ^ (lane & 7)) << 4)) =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
      __syncwarp();
      {
        unsigned char* const gbase = reinterpret_cast<unsigned char*>(e.z_out + ((size_t)tile * TM + q * 32) * C_Z);
        const int rr = lane >> 4, ch = lane & 15;
#pragma unroll
        for (int i16 = 0; i16 < 16; ++i16) {
          const int row = i16 * 2 + rr;
          const uint4 pk = *reinterpret_cast<const uint4*>(obuf + row * 256 + ((ch ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(gbase + row * 256 + ch * 16) = pk;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

void edge_embed_tc2(const EdgeEmbedArgs& a, cudaStream_t st) {
  S2S_CHECK(((size_t)a.B * a.L * a.L) % TM == 0, "edge_embed_tc2 needs B*L*L % 128 == 0 (api.cu pads chain lengths to a multiple of 32)");
  const bool flat = a.L % TM != 0;
  S2S_CHECK(a.wimg, "edge_embed_tc2: weight image missing");
  EePipeArgs k;
  k.e = a; k.wimg = a.wimg; k.n_tiles = (int)((size_t)a.B * a.L * a.L / TM);
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    configured = true;
  }
  S2S_CHECK(a.vec4, "edge_embed_tc2: packed bias / LayerNorm vector missing");
  S2S_CUDA(cudaMemcpyToSymbolAsync(c_ee, a.vec4, sizeof(float) * 4 * C_Z, 0, cudaMemcpyDeviceToDevice, st));
  S2S_PROF("edge_embed", st);
  const int cap = sm_count();
  if (flat) edge_embed_pipe_kernel<true><<<k.n_tiles < cap ? k.n_tiles : cap, P_THREADS, P_SMEM, st>>>(k);
  else edge_embed_pipe_kernel<false><<<k.n_tiles < cap ? k.n_tiles : cap, P_THREADS, P_SMEM, st>>>(k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
