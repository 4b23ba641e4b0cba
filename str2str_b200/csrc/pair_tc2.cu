// Fused EdgeTransition, second generation: MMA / epilogue overlap and BOTH hidden activations resident in tensor memory.
//
// Same math, tile shape (128 pair rows) and operand conventions as pair_tc.cu.  What changed, and why (measured on B200,
// see profiles/README.md): the first-generation kernel ran its three layers strictly serially and kept h1/h2 in shared
// memory, which left room for only three 16 KB weight stages.  An ablation (S2S_ET_DEBUG) showed that neither the weight
// stream's bandwidth nor the MMAs were the limit: the 3-deep ring could not cover the L2 latency of a weight block, so
// the tensor pipe idled about half of every block.  Here
//   * h1 AND h2 live in tensor memory as packed bf16 (written by the epilogue with tcgen05.st, read by tcgen05.mma as
//     the A operand), which frees 96 KB of shared memory: the weight ring is 9 stages deep;
//   * every layer is produced in chunks that alternate between accumulator regions, so the eight epilogue warps drain
//     chunk c while the tensor pipe computes chunk c+1, and the LayerNorm/store epilogue overlaps the next tile's layer 1;
//   * A-from-TMEM MMAs also halve the shared-memory read traffic (an SS-mode 128x128x16 MMA reads 8 KB per 64 cycles,
//     exactly the 128 B/clk the shared memory can deliver).
// TMEM map (512 columns x 128 lanes):
//   [  0,128)  layer-1 accumulator 0  | layer-2 accumulators (2 x 64 columns) | final-layer accumulator
//   [128,256)  layer-1 accumulator 1  | then h2 columns   0..255 (packed bf16, 4 x 32 columns)
//   [256,448)  h1 (384 packed bf16)
//   [448,512)  h2 columns 256..383
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
constexpr int ET2_THREADS = 64 + 32 * NEW;
constexpr int VEC_FLOATS = D_ET + C_Z + D_ET + C_Z + C_Z + 2 * NPART * 128;  // u_i, p_i, b2, ln_w, ln_b, LayerNorm partial sums
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
constexpr int N_BARS = 2 * NSTAGE + 12;
constexpr int SMEM_BYTES = OFF_BAR + N_BARS * 8 + 16;

constexpr uint32_t COL_H1 = 256;
__device__ __forceinline__ uint32_t h2_col(int chunk) { return chunk < 4 ? 128u + 32u * chunk : 448u + 32u * (chunk - 4); }

struct Args {
  const bf16* wimg;
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles, ncopy;
  int dbg;  // timing experiments only (S2S_ET_DEBUG): 1 no weight TMA, 2 no MMA, 4 no epilogue math
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

template <bool MC>
__global__ void __launch_bounds__(ET2_THREADS, 1)
edge_transition_tc2_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n, Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B operand blocks need 1024-byte alignment
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* u_s = reinterpret_cast<float*>(smem + OFF_VEC);
  float* p_s = u_s + D_ET;
  float* b2_s = p_s + C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  float* red_s = lnb_s + C_Z;  // [2 stats][NPART column parts][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NSTAGE;
  uint64_t* a0_full = bars + 2 * NSTAGE;
  uint64_t* a0_empty = a0_full + 1;
  // accumulator regions: R0 = cols [0,64), R1 = [64,128), R2 = [128,256).  Uses: "E" = R0+R1 as one 128-column
  // accumulator (layer-1 chunks 0 and 2, final layer; drained by all 8 epilogue warps), R2 (layer-1 chunk 1), and
  // "G0"/"G1" = R0 / R1 alone (even / odd layer-2 chunks; drained by epilogue group 0 / 1 only).
  uint64_t* fullE = a0_full + 2;
  uint64_t* full2 = a0_full + 3;
  uint64_t* fullG = a0_full + 4;    // [2]
  uint64_t* emptyE = a0_full + 6;
  uint64_t* empty2 = a0_full + 7;
  uint64_t* emptyG = a0_full + 8;   // [2]
  uint64_t* h1_full = a0_full + 10;
  uint64_t* h2_full = a0_full + 11;
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
    mbar_init(&fullG[0], 1);
    mbar_init(&fullG[1], 1);
    mbar_init(emptyE, 32 * NEW);
    mbar_init(empty2, 32 * NEW);
    mbar_init(&emptyG[0], 16 * NEW);
    mbar_init(&emptyG[1], 16 * NEW);
    mbar_init(h1_full, 32 * NEW);
    mbar_init(h2_full, 32 * NEW);
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
  const int tiles_per_i = a.L / TM;
  constexpr uint32_t IDESC128 = make_idesc(128, 128), IDESC64 = make_idesc(128, 64);
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t cnt = 0, ph_a0 = 0;
      const bf16* wimg = a.wimg + (size_t)(blockIdx.x % a.ncopy) * ((size_t)WTILES * TILE_BYTES / 2);
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
        const int b = bi / a.L;
        mbar_wait(a0_empty, ph_a0 ^ 1);
        ph_a0 ^= 1;
        mbar_expect_tx(a0_full, 4 * TILE_BYTES);
        tma_load_2d(smem + OFF_A0, &tmap_z, 0, tile * TM, a0_full);
        tma_load_2d(smem + OFF_A0 + TILE_BYTES, &tmap_z, KBLK, tile * TM, a0_full);
        tma_load_2d(smem + OFF_A0 + 2 * TILE_BYTES, &tmap_n, 0, b * a.L + j0, a0_full);
        tma_load_2d(smem + OFF_A0 + 3 * TILE_BYTES, &tmap_n, KBLK, b * a.L + j0, a0_full);
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
      uint32_t nE = 0, n2 = 0, nG0 = 0, nG1 = 0;  // barrier phase bookkeeping (waits issued so far)
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
        for (int nc = 0; nc < 3; ++nc) {
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
        // ---- layer 2: six 64-column chunks alternating R0 / R1; A = h1 from tensor memory ----
        mbar_wait(h1_full, ph_h1);
        ph_h1 ^= 1;
        tc_fence_after();
        wait_prev(emptyE, nE);  // layer-1 chunk 2 has left R0+R1
        for (int c = 0; c < 6; ++c) {
          if (c >= 2) { if (c & 1) wait_done(&emptyG[1], nG1); else wait_done(&emptyG[0], nG0); }
          tc_fence_after();
          const uint32_t d = tmem + (c & 1) * 64;
          for (int kp = 0; kp < 3; ++kp, ++cnt) {  // one block = [64 n x 128 k]: two 64-row K-blocks of 8 KB
            const uint32_t wb = next_block();
            if (elect_one()) {
              if (do_mma) {
                kblock_ts(d, tmem + COL_H1 + kp * 64, wb, IDESC64, kp == 0);
                kblock_ts(d, tmem + COL_H1 + kp * 64 + 32, wb + BLK / 2, IDESC64, false);
              }
              release_block();
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&fullG[c & 1]);
          __syncwarp();
        }
        // ---- final layer into R0+R1: [z | n'_j] terms (A in shared memory), then h2 (A in tensor memory) ----
        wait_done(&emptyG[0], nG0);  // layer-2 chunks 4 and 5 have left R0 / R1
        wait_done(&emptyG[1], nG1);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t wb = next_block();
          if (elect_one()) {
            if (do_mma) kblock_ss(tmem, a0 + kb * BLK, wb, IDESC128, kb == 0);
            release_block();
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(a0_empty);  // the activation tile is free: the next tile's TMA overlaps the rest of this layer
        __syncwarp();
        mbar_wait(h2_full, ph_h2);
        ph_h2 ^= 1;
        tc_fence_after();
        for (int c = 0; c < 6; ++c, ++cnt) {
          const uint32_t wb = next_block();
          if (elect_one()) {
            if (do_mma) kblock_ts(tmem, tmem + h2_col(c), wb, IDESC128, false);
            release_block();
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(fullE);
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
    constexpr int GW = NPART / 2;        // warps per quarter in one layer-2 group
    const int grp = part / GW, sub = part % GW;  // layer 2: group grp drains chunks grp, grp+2, grp+4; sub splits the 64 columns
    constexpr int CW2 = 64 / GW;
    uint32_t fE = 0, f2 = 0, fG = 0;  // completed uses seen per "full" barrier
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / a.L;
      named_bar_sync(1, 32 * NEW);
      for (int c = et; c < D_ET; c += 32 * NEW) u_s[c] = a.u[(size_t)bi * D_ET + c];
      if (et < C_Z) p_s[et] = a.p[(size_t)bi * C_Z + et];
      named_bar_sync(1, 32 * NEW);
      const float m = a.mask[bi] * a.mask[(size_t)b * a.L + j0 + r];
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
          add_vec(u_s + nc * 128 + part * CW, CW);
          store_packed(COL_H1 + nc * 64 + part * (CW / 2), CW);
        }
        tc_fence_before();
        mbar_arrive(nc == 1 ? empty2 : emptyE);
      }
      mbar_arrive(h1_full);
      // ---- layer 2: + b2, relu, pack, into tensor memory as the final layer's A operand.  The two epilogue groups
      //      take the even / odd chunks, so each has two chunk-MMA times to turn one chunk around ----
      for (int c = grp; c < 6; c += 2) {
        wait_full(&fullG[grp], fG);
        if (do_epi) {
          load_cols(grp * 64 + sub * CW2, CW2);
          add_vec(b2_s + c * 64 + sub * CW2, CW2);
          store_packed(h2_col(c) + sub * (CW2 / 2), CW2);
        }
        tc_fence_before();
        mbar_arrive(&emptyG[grp]);
      }
      mbar_arrive(h2_full);
      // ---- output: + p_i, LayerNorm over 128 channels (exact two-pass; the NPART threads of a row exchange partial sums
      //      through shared memory), * edge mask, bf16 store ----
      {
        wait_full(fullE, fE);
        if (do_epi) load_cols(part * CW, CW);
        tc_fence_before();
        mbar_arrive(emptyE);  // accumulator is in registers: the tensor pipe may reuse it
        if (!do_epi) continue;
        add_vec(p_s + part * CW, CW);
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

size_t et2_wimg_elems() { return (size_t)WTILES * TM * KBLK; }

// Weight blocks (16 KB each) in exactly the order the MMA issuer consumes them.
void build_et2_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  auto block = [&](const float* src, int n0, int k0, int rows) {
    build_wblock_kernel<<<ceil_div(rows * KBLK, 256), 256, 0, st>>>(src, D_ET, n0, k0, rows, d);
    S2S_LAUNCH_CHECK();
    d += rows * KBLK * 2;
  };
  auto aug = [](int kb) { return kb < 2 ? kb * KBLK : 256 + (kb - 2) * KBLK; };  // [z | n'_j] columns of a 384-wide weight
  for (int nc = 0; nc < 3; ++nc)  // layer 1: [128 n x 64 k]
    for (int kb = 0; kb < 4; ++kb) block(W1, nc * 128, aug(kb), 128);
  for (int c = 0; c < 6; ++c)     // layer 2: [64 n x 128 k] as two 64-row K-blocks
    for (int kp = 0; kp < 3; ++kp) {
      block(W2, c * 64, (2 * kp) * KBLK, 64);
      block(W2, c * 64, (2 * kp + 1) * KBLK, 64);
    }
  for (int kb = 0; kb < 4; ++kb) block(Wf, 0, aug(kb), 128);   // final: [z | n'_j] terms
  for (int kb = 0; kb < 6; ++kb) block(Wf, 0, kb * KBLK, 128);  // final: h2 terms
}

void edge_transition_tc2(const EdgeTransitionArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % TM == 0, "edge_transition_tc2 needs L % 128 == 0");
  S2S_CHECK(a.wimg2 && a.nprime_bf16, "edge_transition_tc2: weight image / bf16 node embedding missing");
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z_in, rows, C_Z, C_Z);
  const CUtensorMap mn = make_bf16_2d_map(a.nprime_bf16, (size_t)a.B * a.L, C_Z, C_Z);
  Args k;
  k.wimg = a.wimg2; k.u = a.u; k.p = a.p; k.b2 = a.b2; k.ln_w = a.ln_w; k.ln_b = a.ln_b; k.mask = a.mask;
  k.z_out = a.z_out; k.L = a.L; k.n_tiles = (int)(rows / TM); k.ncopy = a.wimg_copies;
  {
    const char* e = getenv("S2S_ET_DEBUG");
    k.dbg = e ? atoi(e) : 0;
  }
  static bool configured = false;
  const int smem = SMEM_BYTES + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  static const int mc_env = [] { const char* e = getenv("S2S_ET_MULTICAST"); return e ? atoi(e) : 1; }();
  const bool mc = mc_env && grid >= 2 && k.n_tiles % 2 == 0;
  if (mc) {
    grid &= ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ET2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // persistent kernel: never launch more clusters than can be co-resident (GPCs with an odd SM count strand one SM)
    static int max_clusters = 0;
    if (!max_clusters) {
      S2S_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, edge_transition_tc2_kernel<true>, &cfg));
      S2S_CHECK(max_clusters > 0, "edge_transition_tc2: no 2-CTA cluster fits");
    }
    if (grid > 2 * max_clusters) grid = 2 * max_clusters;
    cfg.gridDim = dim3(grid);
    S2S_CUDA(cudaLaunchKernelEx(&cfg, edge_transition_tc2_kernel<true>, mz, mn, k));
  } else {
    edge_transition_tc2_kernel<false><<<grid, ET2_THREADS, smem, st>>>(mz, mn, k);
  }
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
