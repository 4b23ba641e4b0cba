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
// Tiles are 128 consecutive rows of the flattened pair tensor [B*L*L][128].  MODE 0 (L % 128 == 0): a tile lies in one
// (decoy, i) row.  MODE 1 (any other L with B*L*L % 128 == 0; api.cu pads chains to a multiple of 32): every row
// carries its own (decoy, i, j), so short chains fill their tiles with several i rows.
//
// MODE 2 — the embedding TABLE.  The layer-1 input of a pair row is t_i | t_j | pos(idx_i - idx_j) | one-hot distogram bin
// (denoising_ipa.py:126-158), and t_i = [time embedding of the decoy | fixed_i]: apart from the fixed flag nothing in a pair
// row's input depends on (i, j) except the index offset and the bin.  So inside one decoy the whole 3-layer MLP + LayerNorm
// is a function of (fixed_i, fixed_j, offset, bin): n_off x 23 distinct rows (x 4 when the decoy has fixed residues)
// instead of L^2 — 11 753 instead of 65 536 at L = 256, 23 529 instead of 262 144 at L = 512.  MODE 2 runs the same pipeline
// over those table rows (the G station builds layer 1 from the row's (offset, bin) instead of from a pair), and
// edge_embed_expand_kernel then writes z by copying each pair's table row: the embedder becomes an HBM-write-bound copy
// plus 1/5 - 1/11 of the matrix work.  Bit-identical to MODE 0 / 1: a table row is the same fp32 sum (t_i + t_j) + pos (+ bin
// row) and goes through the same MMAs, statistics and bf16 rounding as the pair rows it stands for.  A setup kernel groups
// each decoy's residues by their fixed value; inputs the table cannot represent (more than two distinct fixed values in a
// decoy, masks other than 0 / 1) raise a device flag that turns the table kernels into no-ops and lets the direct kernel,
// launched after them, do the work instead.
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
  // table mode (MODE 2) / fallback control: ctl[0] = active table blocks, ctl[1] = "table not applicable" flag,
  // ctl[2], ctl[3] = setup scratch (zero between launches), ctl[4 ..] = active (decoy * 4 + variant) blocks, then rep[B][2] = representative residue row per fixed class (-1: none)
  const int* ctl = nullptr;
  int only_if_flag = 0;  // MODE 0 / 1 launched behind the table kernels: run only when ctl[1] is set
};

constexpr int TAB_SLOTS = N_BINS + 1;  // no bin, bin 0 .. 21
__host__ __device__ inline int tab_rows_per_block(int n_off) { return (n_off * TAB_SLOTS + TM - 1) / TM * TM; }

template <int MODE>
__global__ void __launch_bounds__(P_THREADS, 1) edge_embed_pipe_kernel(EePipeArgs a) {
  constexpr bool FLAT = MODE == 1, TABLE = MODE == 2;
  if (a.only_if_flag && a.ctl[1] == 0) return;  // the table kernels did the work
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
  pdl_sync();  // only weights (constant during an iteration) were read so far
  const int tiles_per_i = (FLAT || TABLE) ? 1 : e.L / TM;
  const int rpb = tab_rows_per_block(e.n_off), tiles_per_blk = rpb / TM;  // table mode: rows / tiles per (decoy, variant) block
  const int* const blk_list = a.ctl + 4;
  const int* const rep = a.ctl + 4 + 4 * e.B;
  if constexpr (TABLE) a.n_tiles = a.ctl[0] * tiles_per_blk;  // 0 when the flag is raised
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
      if constexpr (TABLE) {
        // table row idx of block (decoy b, variant v): offset = idx / 23, slot = idx % 23 (0 = no bin, k + 1 = bin k)
        const int blk = tile / tiles_per_blk, idx0 = (tile - blk * tiles_per_blk) * TM + r_begin;
        const int bv = blk_list[blk], b = bv >> 2;
        const int ri = rep[2 * b + ((bv >> 1) & 1)], rj = rep[2 * b + (bv & 1)];
        const int idx_l = idx0 + (lane < r_cnt ? lane : 0);
        const bool ok_l = idx_l < e.n_off * TAB_SLOTS;
        const int off_l = ok_l ? idx_l / TAB_SLOTS : 0;
        const int bin_l = ok_l ? idx_l - off_l * TAB_SLOTS - 1 : -1;
        const float4 tiv = __ldg(reinterpret_cast<const float4*>(e.Ti + (size_t)ri * C_Z + c));
        const float4 tjv = __ldg(reinterpret_cast<const float4*>(e.Tj + (size_t)rj * C_Z + c));
        mbar_wait(&a1_empty[s], ph ^ 1);
#pragma unroll 6
        for (int r16 = 0; r16 < r_cnt; ++r16) {
          const int bin = __shfl_sync(0xffffffffu, bin_l, r16);
          const int off = __shfl_sync(0xffffffffu, off_l, r16);
          const float4 tpv = __ldg(reinterpret_cast<const float4*>(e.Tpos + (size_t)off * C_Z + c));
          float4 h = make_float4(tiv.x + tjv.x + tpv.x, tiv.y + tjv.y + tpv.y, tiv.z + tjv.z + tpv.z, tiv.w + tjv.w + tpv.w);
          if (bin >= 0) {
            const float4 wv = *reinterpret_cast<const float4*>(wd_s + bin * P_WD_PITCH + c);
            h.x += wv.x; h.y += wv.y; h.z += wv.z; h.w += wv.w;
          }
          *reinterpret_cast<uint2*>(dst + sw128_offset(r_begin + r16, c % KBLK)) =
              make_uint2(pack_bf16(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f)), pack_bf16(fmaxf(h.z, 0.f), fmaxf(h.w, 0.f)));
        }
      } else if constexpr (FLAT) {
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
      float m = 1.f;  // table rows are stored unmasked: the expand kernel writes zeros for masked pairs
      if constexpr (!TABLE) {
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
        m = __ldg(e.mask + bi) * __ldg(e.mask + bj);
      }
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
        size_t row0 = (size_t)tile * TM + q * 32;
        if constexpr (TABLE) {  // block (b, v) of the table lives at row (b * 4 + v) * rpb whether or not earlier blocks are active
          const int blk = tile / tiles_per_blk;
          row0 = (size_t)blk_list[blk] * rpb + (size_t)(tile - blk * tiles_per_blk) * TM + q * 32;
        }
        unsigned char* const gbase = reinterpret_cast<unsigned char*>((TABLE ? e.table : e.z_out) + row0 * C_Z);
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

// Groups every decoy's residues by their fixed value (class 0 = the value of residue 0, class 1 = the other value) and lists
// the table blocks to build.  One CTA per decoy scans its residues; the last CTA to finish writes the block list.
__global__ void __launch_bounds__(256) embed_table_setup_kernel(int B, int L, const float* __restrict__ fixed, const float* __restrict__ mask,
                                                                unsigned char* __restrict__ cls, int* __restrict__ ctl, int* __restrict__ sync_ws,
                                                                int variants) {
  pdl_sync();
  __shared__ int rep1_s, bad_s, last_s;
  const int b = blockIdx.x;
  int* const rep = ctl + 4 + 4 * B;
  if (threadIdx.x == 0) { rep1_s = L; bad_s = 0; }
  __syncthreads();
  const float f0 = fixed[(size_t)b * L];
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float f = fixed[(size_t)b * L + i], m = mask[(size_t)b * L + i];
    if (m != 0.f && m != 1.f) bad_s = 1;
    const int c = f != f0;
    cls[(size_t)b * L + i] = (unsigned char)c;
    if (c) atomicMin(&rep1_s, i);
  }
  __syncthreads();
  const int rep1 = rep1_s;
  if (rep1 < L) {
    const float f1 = fixed[(size_t)b * L + rep1];
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const float f = fixed[(size_t)b * L + i];
      if (f != f0 && f != f1) bad_s = 1;  // a third distinct value (or NaN)
    }
    if (variants < 4) bad_s = 1;          // the host planned a table without fixed-residue variants
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    rep[2 * b] = b * L;
    rep[2 * b + 1] = rep1 < L ? b * L + rep1 : -1;
    if (bad_s) atomicOr(&sync_ws[1], 1);
    __threadfence();
    last_s = atomicAdd(&sync_ws[0], 1) == B - 1;
  }
  __syncthreads();
  if (last_s && threadIdx.x == 0) {  // every decoy's rep[] is visible (fence + counter): build the list, reset the scratch
    __threadfence();
    int n = 0;
    for (int d = 0; d < B; ++d) {
      ctl[4 + n++] = d * 4;
      if (((volatile int*)rep)[2 * d + 1] >= 0)
        for (int v = 1; v < 4; ++v) ctl[4 + n++] = d * 4 + v;
    }
    const int bad = ((volatile int*)sync_ws)[1];
    ctl[0] = bad ? 0 : n;
    ctl[1] = bad;
    sync_ws[0] = 0;
    sync_ws[1] = 0;
  }
}

// z[b,i,j,:] = table row of (class_i, class_j, idx_i - idx_j, bin(|ca_i - ca_j|)), or zeros for masked pairs.  A warp takes 32
// consecutive pair rows: lanes classify one row each, then the warp copies the rows two at a time (half a warp per 256-byte
// row), all sixteen 16-byte loads of a lane in flight at once.  Reads come from L2 (a decoy's table is ~3 MB), writes are the
// 1 GB of z: the kernel is bound by the HBM write.
__global__ void __launch_bounds__(256) edge_embed_expand_kernel(EePipeArgs a, long n_rows) {
  pdl_sync();
  const EdgeEmbedArgs& e = a.e;
  if (a.ctl[1]) return;  // table not applicable: the direct kernel runs instead
  __shared__ float edge_s[N_BINS];
  if (threadIdx.x < N_BINS) edge_s[threadIdx.x] = e.bin_lower[threadIdx.x];
  __syncthreads();
  const float inv_step = (float)(N_BINS - 1) / (edge_s[N_BINS - 1] - edge_s[0]);
  const int lane = threadIdx.x & 31;
  const int rpb = tab_rows_per_block(e.n_off);
  const long warp_id = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long)gridDim.x * blockDim.x) >> 5;
  const uint4* const tab = reinterpret_cast<const uint4*>(e.table);
  uint4* const zo = reinterpret_cast<uint4*>(e.z_out);
  for (long r0 = warp_id * 32; r0 < n_rows; r0 += n_warps * 32) {
    const long f = r0 + lane;  // n_rows % 32 == 0
    const int bi = (int)(f / e.L), j = (int)(f - (long)bi * e.L), b = bi / e.L;
    const int bj = b * e.L + j;
    long src_l = -1;
    if (__ldg(e.mask + bi) * __ldg(e.mask + bj) != 0.f) {
      const int bin = pair_distogram_bin_fast(e.sc_ca + (size_t)bi * 3, e.sc_ca + (size_t)bj * 3, edge_s, inv_step);
      const int off = min(max((int)(e.ridx[bi] - e.ridx[bj]) - e.d_min, 0), e.n_off - 1);
      const int v = e.cls[bi] * 2 + e.cls[bj];
      src_l = ((long)(b * 4 + v) * rpb + off * TAB_SLOTS + bin + 1) * 16;  // in 16-byte pieces
    }
    const int half = lane >> 4, piece = lane & 15;
    uint4 v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const long src = __shfl_sync(0xffffffffu, src_l, 2 * k + half);
      v[k] = src >= 0 ? __ldg(tab + src + piece) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) zo[(r0 + 2 * k + half) * 16 + piece] = v[k];
  }
}

}  // namespace

size_t edge_embed_table_elems(int B, int n_off, int variants) { return (size_t)B * 4 * tab_rows_per_block(n_off) * C_Z * (variants >= 1 ? 1 : 0); }
size_t edge_embed_ctl_ints(int B) { return 4 + 4 * (size_t)B + 2 * (size_t)B; }

void edge_embed_tc2(const EdgeEmbedArgs& a, cudaStream_t st) {
  S2S_CHECK(((size_t)a.B * a.L * a.L) % TM == 0, "edge_embed_tc2 needs B*L*L % 128 == 0 (api.cu pads chain lengths to a multiple of 32)");
  S2S_CHECK(a.wimg, "edge_embed_tc2: weight image missing");
  const bool flat = a.L % TM != 0;
  EePipeArgs k;
  k.e = a; k.wimg = a.wimg; k.n_tiles = (int)((size_t)a.B * a.L * a.L / TM);
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    configured = true;
  }
  S2S_CHECK(a.vec4, "edge_embed_tc2: packed bias / LayerNorm vector missing");
  S2S_CUDA(cudaMemcpyToSymbolAsync(c_ee, a.vec4, sizeof(float) * 4 * C_Z, 0, cudaMemcpyDeviceToDevice, st));
  S2S_PROF("edge_embed", st);
  const int cap = sm_count();
  if (a.table_variants > 0) {
    S2S_CHECK(a.table && a.tab_ctl && a.cls, "edge_embed_tc2: table buffers missing");
    k.ctl = a.tab_ctl;
    launch_pdl(embed_table_setup_kernel, a.B, 256, 0, st, a.B, a.L, a.fixed, a.mask, a.cls, a.tab_ctl, a.tab_ctl + 2, a.table_variants);
    S2S_LAUNCH_CHECK();
    const int max_tiles = a.B * a.table_variants * (tab_rows_per_block(a.n_off) / TM);
    launch_pdl(edge_embed_pipe_kernel<2>, max_tiles < cap ? max_tiles : cap, P_THREADS, P_SMEM, st, k);
    S2S_LAUNCH_CHECK();
    const long n_rows = (long)a.B * a.L * a.L;
    launch_pdl(edge_embed_expand_kernel, cap * 8, 256, 0, st, k, n_rows);
    S2S_LAUNCH_CHECK();
    k.only_if_flag = 1;  // the direct kernel below runs only if the setup kernel found inputs the table cannot represent
  }
  if (flat) launch_pdl(edge_embed_pipe_kernel<1>, k.n_tiles < cap ? k.n_tiles : cap, P_THREADS, P_SMEM, st, k);
  else launch_pdl(edge_embed_pipe_kernel<0>, k.n_tiles < cap ? k.n_tiles : cap, P_THREADS, P_SMEM, st, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
