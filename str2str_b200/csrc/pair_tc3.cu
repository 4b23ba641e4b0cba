// Fused EdgeTransition, third generation: the second-generation kernel (pair_tc2.cu) with layer 2 issued as three
// 128-column chunks instead of six 64-column ones.
//
// Why: a TS-mode MMA re-reads its 128 x 16 A operand from tensor memory for every instruction, so an N = 64 chunk pays the
// same A traffic as an N = 128 one for half the work, and tensor-memory reads are also what the epilogue warps need
// (tcgen05.ld) — measured with S2S_ET_DEBUG on the second-generation kernel: issuing layer 2 with N = 128 instructions
// takes the MMA-only time from 2.12 to 1.93 ms and the full kernel from 2.98 to 2.78 ms.  Two 128-column accumulators do not
// fit next to h1 (192 columns) and h2 (192 columns), so every h2 chunk is stored IN PLACE over the first half of the
// accumulator it was drained from (or into the one spare 64-column strip), and the final layer accumulates in the
// then-dead h1 region.  TMEM map (512 columns x 128 lanes):
//   E  = [  0,128)  accumulator: layer-1 chunks 0, 2; layer-2 chunk 1         -> afterwards h2 chunk 1 in [0,64)
//   R2 = [128,256)  accumulator: layer-1 chunk 1;     layer-2 chunks 0, 2      -> afterwards h2 chunk 2 in [128,192)
//   H1 = [256,448)  h1 (384 packed bf16)                                      -> afterwards the final-layer accumulator [256,384)
//   S  = [448,512)  h2 chunk 0
// Ordering that makes the aliasing safe: tcgen05.mma instructions of one thread execute in issue order, so the final
// layer's reads of h2 precede the next tile's layer-1 writes into E / R2, and its accumulation into the h1 region
// follows the last layer-2 MMA's reads of h1; epilogue threads of one row quarter synchronise before an in-place store
// (other column parts of the same rows are read by other warps) and once per tile before h1 is rewritten.
//
// Row tiles are 128 consecutive rows of the FLATTENED pair tensor [B*L*L][128].  With L % 128 == 0 a tile lies inside one
// (decoy, i) row (FLAT = false: one u_i / p_i vector and one 128-row n'_j box per tile).  For any other L % 32 == 0
// (FLAT = true; the library pads chain lengths to a multiple of 32, api.cu) a tile is four 32-row segments, each inside one
// (decoy, i): the n'_j rows arrive as four 32-row TMA boxes, and every TMEM lane quarter (= one segment) adds its own
// u_i / p_i vector, so L = 64 fills its tiles with two i rows and L = 96 / 160 / 320 ... run at the same per-row rate.
#include <cstdlib>

#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int NSTAGE = 9;
constexpr int WTILES = 40;                  // 12 (layer 1) + 18 (layer 2) + 10 (final) blocks of 16 KB per row tile
constexpr int OFF_A0 = 0;                   // [z | n'_j] tile: 4 K-blocks
constexpr int OFF_W = 4 * TILE_BYTES;       // weight ring
constexpr int OFF_VEC = OFF_W + NSTAGE * TILE_BYTES;
constexpr int NEW = 16;                     // epilogue warps: NEW/4 per TMEM lane quarter, each owning 128/(NEW/4) columns of a 128-column chunk
constexpr int NPART = NEW / 4;
constexpr int CW = 128 / NPART;             // accumulator columns per thread per 128-column chunk
constexpr int ET3_THREADS = 64 + 32 * NEW;
constexpr int NSEG = 4;                     // 32-row segments of a tile (FLAT: each has its own (decoy, i))
constexpr int VEC_FLOATS = NSEG * (D_ET + C_Z) + D_ET + C_Z + C_Z + 2 * NPART * 128;  // u_i, p_i per segment, b2, ln_w, ln_b, LayerNorm partial sums
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
constexpr int N_BARS = 2 * NSTAGE + 18;
constexpr int SMEM_BYTES = OFF_BAR + N_BARS * 8 + 16;

constexpr uint32_t COL_H1 = 256;
constexpr uint32_t COL_FIN = 256;  // final-layer accumulator (over the dead h1)
// packed h2 columns of K-block kb (64 k = 32 columns): chunk kb/2 lives in S, R2[0:64), E[0:64)
__device__ __forceinline__ uint32_t h2_col(int kb) { return (kb < 2 ? 448u : kb < 4 ? 0u : 128u) + 32u * (kb & 1); }
__device__ __forceinline__ uint32_t h2_chunk_col(int c) { return c == 0 ? 448u : c == 1 ? 0u : 128u; }

struct Args {
  const bf16* wimg;
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles, ncopy;
  int dbg;  // timing experiments only (S2S_ET_DEBUG): 1 no weight TMA, 2 no MMA, 4 no epilogue math
  int interleave;  // layer-1 chunks 0 and 1 issued interleaved (two independent accumulator chains)
};

// MC = true: the kernel runs as clusters of two CTAs that share the weight stream: each 16 KB weight block is fetched from
// L2 once per PAIR (the CTAs alternate as loader) and TMA-multicast into both rings, so the L2 -> SM weight traffic — the
// resource this kernel saturates, ~8 TB/s measured — halves.  A ring slot is released by both CTAs' MMA issuers
// (multicast tcgen05.commit, barrier count 2).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_1d_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <bool MC, bool FLAT>
__global__ void __launch_bounds__(ET3_THREADS, 1)
edge_transition_tc3_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n, Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B operand blocks need 1024-byte alignment
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* u_s = reinterpret_cast<float*>(smem + OFF_VEC);  // [NSEG][D_ET]
  float* p_s = u_s + NSEG * D_ET;                         // [NSEG][C_Z]
  float* b2_s = p_s + NSEG * C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  float* red_s = lnb_s + C_Z;  // [2 stats][NPART column parts][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NSTAGE;
  uint64_t* a0_full = bars + 2 * NSTAGE;
  uint64_t* a0_empty = a0_full + 1;
  // accumulators: "E" = columns [0,128), "R2" = [128,256) (alternating chunks of layers 1 and 2), "F" = [256,384) (final layer)
  uint64_t* fullE = a0_full + 2;
  uint64_t* full2 = a0_full + 3;
  uint64_t* fullF = a0_full + 4;    // final-layer accumulator ready
  uint64_t* emptyE = a0_full + 6;
  uint64_t* empty2 = a0_full + 7;
  uint64_t* h1p = a0_full + 12;     // [3] h1 columns of layer-1 chunk nc are in tensor memory (K-blocks 2nc, 2nc+1 of layer 2)
  uint64_t* h2p = a0_full + 15;     // [3] h2 chunk c is in tensor memory (K-blocks 2c, 2c+1 of the final layer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], MC ? 2 : 1);
    }
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    mbar_init(fullE, 1);
    mbar_init(full2, 1);
    mbar_init(fullF, 1);
    mbar_init(emptyE, 32 * NEW);
    mbar_init(empty2, 32 * NEW);
    for (int k = 0; k < 3; ++k) {
      mbar_init(&h1p[k], 32 * NEW);
      mbar_init(&h2p[k], 32 * NEW);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int c = threadIdx.x; c < D_ET; c += blockDim.x) b2_s[c] = a.b2[c];
  for (int c = threadIdx.x; c < C_Z; c += blockDim.x) {
    lnw_s[c] = a.ln_w[c];
    lnb_s[c] = a.ln_b[c];
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // only weights (constant during an iteration) were read so far
  const int tiles_per_i = FLAT ? 1 : a.L / TM;  // FLAT tiles are addressed by flattened row, not by (decoy, i)
  constexpr uint32_t IDESC128 = make_idesc(128, 128), IDESC64 = make_idesc(128, 64);
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t cnt = 0, ph_a0 = 0;
      const bf16* wimg = a.wimg + (size_t)(blockIdx.x % a.ncopy) * ((size_t)WTILES * TILE_BYTES / 2);
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;  // (unused when FLAT)
        const int b = bi / a.L;
        mbar_wait(a0_empty, ph_a0 ^ 1);
        ph_a0 ^= 1;
        mbar_expect_tx(a0_full, 4 * TILE_BYTES);
        tma_load_2d(smem + OFF_A0, &tmap_z, 0, tile * TM, a0_full);
        tma_load_2d(smem + OFF_A0 + TILE_BYTES, &tmap_z, KBLK, tile * TM, a0_full);
        if constexpr (FLAT) {  // tmap_n has 32-row boxes: segment sg holds keys j_sg .. j_sg + 31 of its own decoy
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) {
            const long f = (long)tile * TM + sg * 32;
            const int bi_s = (int)(f / a.L), j_s = (int)(f - (long)bi_s * a.L);
            const int nrow = (bi_s / a.L) * a.L + j_s;
            tma_load_2d(smem + OFF_A0 + 2 * TILE_BYTES + sg * 4096, &tmap_n, 0, nrow, a0_full);
            tma_load_2d(smem + OFF_A0 + 3 * TILE_BYTES + sg * 4096, &tmap_n, KBLK, nrow, a0_full);
          }
        } else {
          tma_load_2d(smem + OFF_A0 + 2 * TILE_BYTES, &tmap_n, 0, b * a.L + j0, a0_full);
          tma_load_2d(smem + OFF_A0 + 3 * TILE_BYTES, &tmap_n, KBLK, b * a.L + j0, a0_full);
        }
        for (int wt = 0; wt < WTILES; ++wt, ++cnt) {
          const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          if (a.dbg & 1) { mbar_arrive(&w_full[s]); continue; }
          mbar_expect_tx(&w_full[s], TILE_BYTES);
          if constexpr (MC) {
            if ((cnt & 1u) == crank)
              tma_bulk_1d_mc(smem + OFF_W + s * TILE_BYTES, wimg + (size_t)wt * (TILE_BYTES / 2), TILE_BYTES, &w_full[s], (uint16_t)3);
          } else {
            tma_bulk_1d(smem + OFF_W + s * TILE_BYTES, wimg + (size_t)wt * (TILE_BYTES / 2), TILE_BYTES, &w_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop (warp-uniform values), one elected lane issues =====
    {
      const uint32_t a0 = desc_lo_sw128(smem_u32(smem + OFF_A0)), wr = desc_lo_sw128(smem_u32(smem + OFF_W));  // descriptor low words
      constexpr uint32_t BLK = TILE_BYTES >> 4;  // one 16 KB block in descriptor address units
      uint32_t cnt = 0, ph_a0 = 0, ph_h1 = 0, ph_h2 = 0;
      uint32_t nE = 0, n2 = 0;  // barrier phase bookkeeping (waits issued so far)
      const bool do_mma = !(a.dbg & 2);
      auto next_block = [&]() -> uint32_t {  // wait for the next streamed weight block
        const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        return wr + s * BLK;
      };
      auto release_block = [&]() {  // called by the elected lane
        if constexpr (MC) umma_commit_mc(&w_empty[cnt % NSTAGE], (uint16_t)3); else umma_commit(&w_empty[cnt % NSTAGE]);
      };
      auto wait_prev = [&](uint64_t* bar, uint32_t& n) {  // wait #k waits for drain #(k-1): the first one passes
        mbar_wait(bar, (n & 1) ^ 1);
        ++n;
      };
      auto wait_done = [&](uint64_t* bar, uint32_t& n) {  // wait #k waits for drain #k
        mbar_wait(bar, n & 1);
        ++n;
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(a0_full, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        // ---- layer 1: three 128-column chunks in R0+R1, R2, R0+R1; A = [z | n'_j] from shared memory ----
        // SS-mode MMAs that accumulate into the SAME tensor-memory columns run as a dependent chain (~260 cycles each in the
        // GEMM experiments, ~190 when consecutive MMAs alternate between two accumulators: profiles/r01c_gemm_and_embedder_
        // experiments.log), so chunks 0 (E) and 1 (R2) are issued interleaved k-step by k-step: block (chunk 0, kb) and block
        // (chunk 1, kb) = ring positions cnt + kb and cnt + 4 + kb are both held, and released together.
        int nc_first = 0;
        if (a.interleave) {
          wait_prev(emptyE, nE);
          wait_prev(empty2, n2);
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t c0 = cnt + kb, c1 = cnt + 4 + kb;
            mbar_wait(&w_full[c0 % NSTAGE], (c0 / NSTAGE) & 1);
            mbar_wait(&w_full[c1 % NSTAGE], (c1 / NSTAGE) & 1);
            tc_fence_after();
            const uint32_t w0 = wr + (c0 % NSTAGE) * BLK, w1 = wr + (c1 % NSTAGE) * BLK, ab = a0 + kb * BLK;
            if (elect_one()) {
              if (do_mma) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (kb | k) { umma_ss<true>(tmem, ab + 2 * k, w0 + 2 * k, IDESC128); umma_ss<true>(tmem + 128, ab + 2 * k, w1 + 2 * k, IDESC128); }
                  else { umma_ss<false>(tmem, ab + 2 * k, w0 + 2 * k, IDESC128); umma_ss<false>(tmem + 128, ab + 2 * k, w1 + 2 * k, IDESC128); }
                }
              }
              if constexpr (MC) { umma_commit_mc(&w_empty[c0 % NSTAGE], (uint16_t)3); umma_commit_mc(&w_empty[c1 % NSTAGE], (uint16_t)3); }
              else { umma_commit(&w_empty[c0 % NSTAGE]); umma_commit(&w_empty[c1 % NSTAGE]); }
            }
            __syncwarp();
          }
          cnt += 8;
          if (elect_one()) { umma_commit(fullE); umma_commit(full2); }
          __syncwarp();
          nc_first = 2;
        }
        for (int nc = nc_first; nc < 3; ++nc) {
          uint32_t d;
          if (nc == 1) { wait_prev(empty2, n2); d = tmem + 128; }
          else { wait_prev(emptyE, nE); d = tmem; }
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb, ++cnt) {
            const uint32_t wb = next_block();
            if (elect_one()) {
              if (do_mma) kblock_ss(d, a0 + kb * BLK, wb, IDESC128, kb == 0);
              release_block();
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(nc == 1 ? full2 : fullE);
          __syncwarp();
        }
        // ---- layer 2: three 128-column chunks in R2, E, R2; A = h1 from tensor memory.  Chunk 0 starts as soon as R2
        //      (layer-1 chunk 1) is drained and consumes h1 K-block by K-block as the layer-1 drains deliver it, so
        //      the tensor pipe does not idle through the drain of layer-1 chunk 2 ----
        for (int c = 0; c < 3; ++c) {
          uint32_t d;
          if (c == 1) { wait_prev(emptyE, nE); d = tmem; }
          else { wait_prev(empty2, n2); d = tmem + 128; }
          tc_fence_after();
          for (int kb = 0; kb < 6; ++kb, ++cnt) {  // one block = [128 n x 64 k]
            if (c == 0 && !(kb & 1)) {
              mbar_wait(&h1p[kb >> 1], ph_h1);
              tc_fence_after();
            }
            const uint32_t wb = next_block();
            if (elect_one()) {
              if (do_mma) kblock_ts(d, tmem + COL_H1 + kb * 32, wb, IDESC128, kb == 0);
              release_block();
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(c == 1 ? fullE : full2);
          __syncwarp();
        }
        ph_h1 ^= 1;
        // ---- final layer into F (the h1 region: its last readers, the layer-2 MMAs above, were issued earlier):
        //      [z | n'_j] terms (A in shared memory), then h2 (A in tensor memory) ----
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t wb = next_block();
          if (elect_one()) {
            if (do_mma) kblock_ss(tmem + COL_FIN, a0 + kb * BLK, wb, IDESC128, kb == 0);
            release_block();
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(a0_empty);  // the activation tile is free: the next tile's TMA overlaps the rest of this layer
        __syncwarp();
        for (int kb = 0; kb < 6; ++kb, ++cnt) {
          if (!(kb & 1)) {
            mbar_wait(&h2p[kb >> 1], ph_h2);
            tc_fence_after();
          }
          const uint32_t wb = next_block();
          if (elect_one()) {
            if (do_mma) kblock_ts(tmem + COL_FIN, tmem + h2_col(kb), wb, IDESC128, false);
            release_block();
          }
          __syncwarp();
        }
        ph_h2 ^= 1;
        if (elect_one()) umma_commit(fullF);
        __syncwarp();
      }
    }
  } else {
    // ===== NEW epilogue warps: TMEM lane quarter = warp % 4; the NPART warps of a quarter split each chunk's columns.
    //       (ablation, S2S_ET_DEBUG: MMAs alone 2.05 ms, epilogue alone 2.2 ms at 8 warps — the epilogue is a chain of
    //        TMEM round trips per warp, so it is bought down with more warps in flight, not with fewer instructions) =====
    const int ew = warp - 2, q = warp & 3, part = ew >> 2, r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool do_epi = !(a.dbg & 4);
    uint32_t fE = 0, f2 = 0, fF = 0;  // completed uses seen per "full" barrier
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      int bi, jr;  // this thread's row: (decoy, i) index and key j
      if constexpr (FLAT) {
        const long f = (long)tile * TM + r;
        bi = (int)(f / a.L);
        jr = (int)(f - (long)bi * a.L);
      } else {
        bi = tile / tiles_per_i;
        jr = (tile % tiles_per_i) * TM + r;
      }
      const int b = bi / a.L;
      named_bar_sync(1, 32 * NEW);
      if constexpr (FLAT) {  // segment sg's vectors: bi of row sg * 32 (uniform over the segment because L % 32 == 0)
        for (int c = et; c < NSEG * D_ET; c += 32 * NEW) {
          const int sg = c / D_ET;
          u_s[c] = a.u[(size_t)(((long)tile * TM + sg * 32) / a.L) * D_ET + (c - sg * D_ET)];
        }
        {
          const int sg = et / C_Z;  // 512 epilogue threads = NSEG * C_Z
          p_s[et] = a.p[(size_t)(((long)tile * TM + sg * 32) / a.L) * C_Z + (et - sg * C_Z)];
        }
      } else {
        for (int c = et; c < D_ET; c += 32 * NEW) u_s[c] = a.u[(size_t)bi * D_ET + c];
        if (et < C_Z) p_s[et] = a.p[(size_t)bi * C_Z + et];
      }
      named_bar_sync(1, 32 * NEW);
      const float* u_q = FLAT ? u_s + q * D_ET : u_s;  // this lane quarter's segment
      const float* p_q = FLAT ? p_s + q * C_Z : p_s;
      const float m = a.mask[bi] * a.mask[(size_t)b * a.L + jr];
      float y[CW];
      auto wait_full = [&](uint64_t* bar, uint32_t& n) {
        mbar_wait(bar, n & 1);
        ++n;
        tc_fence_after();
      };
      auto add_vec = [&](const float* vec, int n) {  // y[0..n) += vec[0..n) (128-bit broadcast shared-memory reads)
#pragma unroll
        for (int e = 0; e < CW; e += 4) {
          if (e < n) {
            const float4 t = *reinterpret_cast<const float4*>(vec + e);
            y[e] += t.x; y[e + 1] += t.y; y[e + 2] += t.z; y[e + 3] += t.w;
          }
        }
      };
      auto load_cols = [&](uint32_t col, int n) {  // y[0..n) <- accumulator columns [col, col+n) of this thread's row
#pragma unroll
        for (int e = 0; e < CW; e += 32)
          if (e < n) tmem_ld32_issue(tmem + lane_off + col + e, y + e);
        tmem_wait_ld();
      };
      auto store_packed = [&](uint32_t col, int n) {  // relu, pack pairs to bf16, into tensor memory as the next A operand
        if constexpr (CW == 64) {
          if (n == 64) {
            uint32_t pk[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) pk[e] = pack_bf16(fmaxf(y[2 * e], 0.f), fmaxf(y[2 * e + 1], 0.f));
            tmem_st32(tmem + lane_off + col, pk);
            return;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(fmaxf(y[2 * e], 0.f), fmaxf(y[2 * e + 1], 0.f));
        tmem_st16(tmem + lane_off + col, pk);
      };
      // ---- layer 1: + u_i, relu, pack, into tensor memory as layer 2's A operand ----
      for (int nc = 0; nc < 3; ++nc) {
        const uint32_t base = nc == 1 ? 128u : 0u;
        if (nc == 1) wait_full(full2, f2); else wait_full(fullE, fE);
        if (do_epi) {
          load_cols(base + part * CW, CW);
          add_vec(u_q + nc * 128 + part * CW, CW);
          store_packed(COL_H1 + nc * 64 + part * (CW / 2), CW);
        }
        tc_fence_before();
        mbar_arrive(nc == 1 ? empty2 : emptyE);
        mbar_arrive(&h1p[nc]);
      }
      // ---- layer 2: + b2, relu, pack, into tensor memory as the final layer's A operand.  Chunk 0 goes to the spare
      //      strip; chunks 1 and 2 overwrite the first half of the accumulator they came from, so the NPART warps of a
      //      row quarter first agree that all of them have read it ----
      for (int c = 0; c < 3; ++c) {
        const uint32_t base = c == 1 ? 0u : 128u;  // accumulators R2, E, R2
        if (c == 1) wait_full(fullE, fE); else wait_full(full2, f2);
        if (do_epi) {
          load_cols(base + part * CW, CW);
          add_vec(b2_s + c * 128 + part * CW, CW);
          if (c > 0) {
            tc_fence_before();
            named_bar_sync(2 + q, 32 * NPART);
            tc_fence_after();
          }
          store_packed(h2_chunk_col(c) + part * (CW / 2), CW);
        }
        tc_fence_before();
        mbar_arrive(c == 1 ? emptyE : empty2);
        mbar_arrive(&h2p[c]);
      }
      // ---- output: + p_i, LayerNorm over 128 channels (exact two-pass; the NPART threads of a row exchange partial sums
      //      through shared memory), * edge mask, bf16 store ----
      {
        wait_full(fullF, fF);
        if (do_epi) load_cols(COL_FIN + part * CW, CW);
        tc_fence_before();
        if (!do_epi) continue;
        add_vec(p_q + part * CW, CW);
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < CW; ++e) sum += y[e];
        red_s[part * 128 + r] = sum;
        named_bar_sync(2 + q, 32 * NPART);
        float tot = 0.f;
#pragma unroll
        for (int pp = 0; pp < NPART; ++pp) tot += red_s[pp * 128 + r];
        const float mean = tot * (1.f / C_Z);
        float sq = 0.f;
#pragma unroll
        for (int e = 0; e < CW; ++e) {
          const float d = y[e] - mean;
          sq += d * d;
        }
        red_s[(NPART + part) * 128 + r] = sq;
        named_bar_sync(2 + q, 32 * NPART);
        float tsq = 0.f;
#pragma unroll
        for (int pp = 0; pp < NPART; ++pp) tsq += red_s[(NPART + pp) * 128 + r];
        const float rstd = rsqrtf(tsq * (1.f / C_Z) + 1e-5f);
        bf16* orow = a.z_out + ((size_t)tile * TM + r) * C_Z + part * CW;
        const float* lw = lnw_s + part * CW;
        const float* lb = lnb_s + part * CW;
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = ((y[c0 + e] - mean) * rstd * lw[c0 + e] + lb[c0 + e]) * m;
          *reinterpret_cast<uint4*>(orow + c0) =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// one [rows x 64] bf16 block in the SW128 K-major layout
__global__ void build_wblock_kernel(const float* __restrict__ src, int ld, int n0, int k0, int rows, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * KBLK) return;
  const int r = idx / KBLK, c = idx % KBLK;
  *reinterpret_cast<bf16*>(dst + sw128_offset(r, c)) = __float2bfloat16_rn(src[(size_t)(n0 + r) * ld + k0 + c]);
}

}  // namespace

size_t et3_wimg_elems() { return (size_t)WTILES * TM * KBLK; }

// Weight blocks (16 KB each) in exactly the order the MMA issuer consumes them.
void build_et3_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  auto block = [&](const float* src, int n0, int k0, int rows) {
    build_wblock_kernel<<<ceil_div(rows * KBLK, 256), 256, 0, st>>>(src, D_ET, n0, k0, rows, d);
    S2S_LAUNCH_CHECK();
    d += rows * KBLK * 2;
  };
  auto aug = [](int kb) { return kb < 2 ? kb * KBLK : 256 + (kb - 2) * KBLK; };  // [z | n'_j] columns of a 384-wide weight
  for (int nc = 0; nc < 3; ++nc)  // layer 1: [128 n x 64 k]
    for (int kb = 0; kb < 4; ++kb) block(W1, nc * 128, aug(kb), 128);
  for (int c = 0; c < 3; ++c)     // layer 2: [128 n x 64 k]
    for (int kb = 0; kb < 6; ++kb) block(W2, c * 128, kb * KBLK, 128);
  for (int kb = 0; kb < 4; ++kb) block(Wf, 0, aug(kb), 128);   // final: [z | n'_j] terms
  for (int kb = 0; kb < 6; ++kb) block(Wf, 0, kb * KBLK, 128);  // final: h2 terms
}

void edge_transition_tc3(const EdgeTransitionArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % 32 == 0, "edge_transition_tc3 needs L % 32 == 0 (api.cu pads chain lengths)");
  static_assert(NEW == 16 && CW == 32, "the in-place h2 stores assume 4 column parts of 32");
  static_assert(32 * NEW == NSEG * C_Z, "p_i staging assumes one element per epilogue thread");
  const bool flat = a.L % TM != 0;
  S2S_CHECK(a.wimg3 && a.nprime_bf16, "edge_transition_tc3: weight image / bf16 node embedding missing");
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z_in, rows, C_Z, C_Z);
  const CUtensorMap mn = make_bf16_2d_map(a.nprime_bf16, (size_t)a.B * a.L, C_Z, C_Z, flat ? 32 : 128);
  Args k;
  k.wimg = a.wimg3; k.u = a.u; k.p = a.p; k.b2 = a.b2; k.ln_w = a.ln_w; k.ln_b = a.ln_b; k.mask = a.mask;
  k.z_out = a.z_out; k.L = a.L; k.n_tiles = (int)(rows / TM); k.ncopy = a.wimg_copies;
  {
    const char* e = getenv("S2S_ET_DEBUG");
    k.dbg = e ? atoi(e) : 0;
    static const int il = [] { const char* v = getenv("S2S_ET_INTERLEAVE"); return v ? atoi(v) : 0; }();
    k.interleave = il;
  }
  static bool configured = false;
  const int smem = SMEM_BYTES + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  static const int mc_env = [] { const char* e = getenv("S2S_ET_MULTICAST"); return e ? atoi(e) : 1; }();
  const bool mc = mc_env && grid >= 2 && k.n_tiles % 2 == 0;
  if (mc) {
    grid &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ET3_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // persistent kernel: never launch more clusters than can be co-resident (GPCs with an odd SM count strand one SM)
    static int max_clusters = 0;
    if (!max_clusters) {
      S2S_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, edge_transition_tc3_kernel<true, false>, &cfg));
      S2S_CHECK(max_clusters > 0, "edge_transition_tc3: no 2-CTA cluster fits");
    }
    if (grid > 2 * max_clusters) grid = 2 * max_clusters;
    cfg.gridDim = dim3(grid);
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see launch_pdl (common.cuh)
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = g_pdl ? 2 : 1;
    if (flat) S2S_CUDA(cudaLaunchKernelEx(&cfg, edge_transition_tc3_kernel<true, true>, mz, mn, k));
    else S2S_CUDA(cudaLaunchKernelEx(&cfg, edge_transition_tc3_kernel<true, false>, mz, mn, k));
  } else if (flat) {
    launch_pdl(edge_transition_tc3_kernel<false, true>, grid, ET3_THREADS, smem, st, mz, mn, k);
  } else {
    launch_pdl(edge_transition_tc3_kernel<false, false>, grid, ET3_THREADS, smem, st, mz, mn, k);
  }
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
