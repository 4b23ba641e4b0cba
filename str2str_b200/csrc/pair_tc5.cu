// Fused EdgeTransition on CTA PAIRS (tcgen05 cta_group::2): the third-generation kernel (pair_tc3.cu: same TMEM map, same
// epilogue, same arithmetic and rounding points) with every MMA issued once per pair of SMs as M = 256.
//
// Why: pair_tc3's ncu capture shows 23.8 GB of L2 -> SM traffic per launch against 2.1 GB of DRAM traffic: every 128-row tile
// pulls all 40 weight blocks (640 KB) through shared memory, and shared-memory bandwidth (operand reads + TMA writes, ~1.6 MB
// per tile) is one of the three resources that kernel runs into.  In a CTA pair the two SMs work on two row
// tiles with ONE weight stream: each CTA stages only half of every weight block (64 of its 128 output columns, 8 KB) and the
// pair's M = 256 MMA reads the B operand half from each SM, so per SM the weight traffic from L2, the shared-memory writes and
// the B-operand reads all halve, and one instruction issue covers two tiles.
//
// Roles per CTA (same as pair_tc3): warp 0 = TMA producer (its own activation tile, its half of the weight stream), warp 1 =
// tensor-memory allocation and, in the LEADER CTA (cluster rank 0), the MMA issuer for the pair; warps 2..17 = epilogue of the
// CTA's own 128 rows.  Everything the issuer waits for lives in the LEADER's shared memory and is signalled by both CTAs:
//   * operand barriers (w_full, a0_full): armed by the leader's producer for the bytes of BOTH CTAs; the peer's TMA loads
//     (cp.async.bulk.tensor ... cta_group::2) complete their bytes on the leader's barrier;
//   * hand-over barriers (h1 / h2 written to tensor memory, accumulators drained): one arrival per epilogue WARP, local in the
//     leader, a remote mbarrier arrive (release.cluster) from the peer.
// Completions of the MMAs go to both CTAs with multicast tcgen05.commit.  (A first cut that forwarded the peer's local
// barriers through a relay warp was bit-exact but 2.4x slower — 6.18 ms vs 2.57 ms at cfg2: every one of the ~56 hand-overs
// per tile pair paid a poll + remote arrive + poll chain on the critical path.)
#include <cstdlib>

#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int NSTAGE = 16;                  // ring of half weight blocks (8 KB each)
constexpr int HALF_BYTES = TILE_BYTES / 2;  // [64 n x 64 k] bf16, SW128 K-major
constexpr int WTILES = 40;                  // 12 (layer 1) + 18 (layer 2) + 10 (final) blocks per row tile
constexpr int OFF_A0 = 0;                   // [z | n'_j] tile: 4 K-blocks
constexpr int OFF_W = 4 * TILE_BYTES;       // weight ring
constexpr int OFF_VEC = OFF_W + NSTAGE * HALF_BYTES;
constexpr int NEW = 16;                     // epilogue warps
constexpr int NPART = NEW / 4;
constexpr int CW = 128 / NPART;
constexpr int ET5_THREADS = 64 + 32 * NEW;
constexpr int NSEG = 4;
constexpr int VEC_FLOATS = 2 * NSEG * (D_ET + C_Z) + D_ET + C_Z + C_Z + 2 * NPART * 128;  // u / p vectors double-buffered by tile parity
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
// barriers: w_full[NSTAGE] w_empty[NSTAGE] a0_full a0_empty fullE full2 fullF emptyE empty2 h1p[3] h2p[3]
// (w_full, a0_full, emptyE, empty2, h1p, h2p are used in the leader CTA only)
constexpr int N_BARS = 2 * NSTAGE + 13;
constexpr int SMEM_BYTES = OFF_BAR + N_BARS * 8 + 16;

constexpr uint32_t COL_H1 = 256;
constexpr uint32_t COL_FIN = 256;
__device__ __forceinline__ uint32_t h2_col(int kb) { return (kb < 2 ? 448u : kb < 4 ? 0u : 128u) + 32u * (kb & 1); }
__device__ __forceinline__ uint32_t h2_chunk_col(int c) { return c == 0 ? 448u : c == 1 ? 0u : 128u; }

struct Args {
  const bf16* wimg;
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles, ncopy;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- cta_group::2 forms of the tcgen05 instructions (one kernel may use one cta_group only) ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of all MMAs issued so far by this thread -> barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
template <bool kAccumulate>
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(kAccumulate ? 1 : 0), "r"(DESC_HI_SW128)
      : "memory");
}
template <bool kAccumulate>
__device__ __forceinline__ void umma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "n"(kAccumulate ? 1 : 0), "r"(DESC_HI_SW128)
      : "memory");
}
__device__ __forceinline__ void kblock2_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool first) {
  if (first) umma2_ss<false>(d, a_lo, b_lo, idesc); else umma2_ss<true>(d, a_lo, b_lo, idesc);
  umma2_ss<true>(d, a_lo + 2, b_lo + 2, idesc);
  umma2_ss<true>(d, a_lo + 4, b_lo + 4, idesc);
  umma2_ss<true>(d, a_lo + 6, b_lo + 6, idesc);
}
__device__ __forceinline__ void kblock2_ts(uint32_t d, uint32_t a_col, uint32_t b_lo, uint32_t idesc, bool first) {
  if (first) umma2_ts<false>(d, a_col, b_lo, idesc); else umma2_ts<true>(d, a_col, b_lo, idesc);
  umma2_ts<true>(d, a_col + 8, b_lo + 2, idesc);
  umma2_ts<true>(d, a_col + 16, b_lo + 4, idesc);
  umma2_ts<true>(d, a_col + 24, b_lo + 6, idesc);
}
// shared::cluster address of the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t cluster_addr(const void* p, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(p)), "r"(cta));
  return raddr;
}
// Remote arrive with RELAXED semantics.  What the issuer needs from a hand-over is ordering of tcgen05 operations (the
// arriving warp's tcgen05.ld / st have completed: tcgen05.wait + tcgen05.fence::before_thread_sync precede the arrive), not
// visibility of generic-proxy memory: the data stays in the peer's tensor memory and is read there by the peer's tensor core.
// A release.cluster arrive compiles to MEMBAR.ALL.GPU + CCTL.IVALL per hand-over and an acquire.cluster wait to an L1
// invalidation per successful poll (measured: 2.99 ms per launch with them, against 2.59 ms for the single-CTA kernel).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// 2-D tensor TMA load of a CTA pair: the data lands in THIS CTA's shared memory, the bytes are completed on the barrier at
// `bar_cluster_addr`, which may belong to the other CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
               : "memory");
}
template <bool FLAT>
__global__ void __launch_bounds__(ET5_THREADS, 1)
edge_transition_pair_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n,
                            const __grid_constant__ CUtensorMap tmap_w, Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* u_s = reinterpret_cast<float*>(smem + OFF_VEC);  // [2][NSEG][D_ET]: per-residue layer-1 terms of this / the next tile
  float* p_s = u_s + 2 * NSEG * D_ET;                     // [2][NSEG][C_Z]: per-residue final-layer terms
  float* b2_s = p_s + 2 * NSEG * C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  float* red_s = lnb_s + C_Z;  // [2 stats][NPART column parts][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NSTAGE;
  uint64_t* a0_full = bars + 2 * NSTAGE;
  uint64_t* a0_empty = a0_full + 1;
  uint64_t* fullE = a0_full + 2;
  uint64_t* full2 = a0_full + 3;
  uint64_t* fullF = a0_full + 4;
  uint64_t* emptyE = a0_full + 5;
  uint64_t* empty2 = a0_full + 6;
  uint64_t* h1p = a0_full + 7;    // [3]
  uint64_t* h2p = a0_full + 10;   // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&w_full[s], 1);   // leader: one arrive.expect_tx covering the half blocks of both CTAs
      mbar_init(&w_empty[s], 1);  // one multicast commit
    }
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    mbar_init(fullE, 1);
    mbar_init(full2, 1);
    mbar_init(fullF, 1);
    mbar_init(emptyE, 2 * NEW);   // leader: one arrival per epilogue warp of the pair
    mbar_init(empty2, 2 * NEW);
    for (int k = 0; k < 3; ++k) {
      mbar_init(&h1p[k], 2 * NEW);
      mbar_init(&h2p[k], 2 * NEW);
    }
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < D_ET; c += blockDim.x) b2_s[c] = a.b2[c];
  for (int c = threadIdx.x; c < C_Z; c += blockDim.x) {
    lnw_s[c] = a.ln_w[c];
    lnb_s[c] = a.ln_b[c];
  }
  __syncthreads();
  cluster_sync_all();                          // both CTAs are resident and their barriers initialised
  if (warp == 1) tmem_alloc2(tmem_slot, 512);  // the same warp of both CTAs, as the 2-SM allocation requires
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // tensor memory is allocated in both CTAs before any MMA of the pair can target it
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // only weights (constant during an iteration) were read so far
  const int tiles_per_i = FLAT ? 1 : a.L / TM;
  constexpr uint32_t IDESC256 = make_idesc(256, 128);  // M = 256 across the pair, N = 128
  // this CTA's tiles: pair p of the grid takes tiles 2p, 2p+1, then strides by the number of pairs
  const int n_pairs_grid = gridDim.x / 2, pair_id = blockIdx.x / 2;
  const int n_pair_tiles = a.n_tiles / 2;

  if (warp == 0) {
    // ===== TMA producer: this CTA's activation tile and its half of every weight block; all bytes complete on the LEADER's
    //       barriers, which the leader's producer arms for both CTAs =====
    if (lane == 0) {
      uint32_t cnt = 0, ph_a0 = 0;
      const int wrow0 = (pair_id % a.ncopy) * (WTILES * TM) + (int)crank * 64;  // row of this CTA's half of block 0 in the weight map
      const uint32_t a0_full_l = cluster_addr(a0_full, 0);
      for (int pt = pair_id; pt < n_pair_tiles; pt += n_pairs_grid) {
        const int tile = 2 * pt + (int)crank;
        const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;  // (unused when FLAT)
        const int b = bi / a.L;
        mbar_wait(a0_empty, ph_a0 ^ 1);
        ph_a0 ^= 1;
        if (leader) mbar_expect_tx(a0_full, 2 * 4 * TILE_BYTES);
        tma_load_2d_pair(smem + OFF_A0, &tmap_z, 0, tile * TM, a0_full_l);
        tma_load_2d_pair(smem + OFF_A0 + TILE_BYTES, &tmap_z, KBLK, tile * TM, a0_full_l);
        if constexpr (FLAT) {
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) {
            const long f = (long)tile * TM + sg * 32;
            const int bi_s = (int)(f / a.L), j_s = (int)(f - (long)bi_s * a.L);
            const int nrow = (bi_s / a.L) * a.L + j_s;
            tma_load_2d_pair(smem + OFF_A0 + 2 * TILE_BYTES + sg * 4096, &tmap_n, 0, nrow, a0_full_l);
            tma_load_2d_pair(smem + OFF_A0 + 3 * TILE_BYTES + sg * 4096, &tmap_n, KBLK, nrow, a0_full_l);
          }
        } else {
          tma_load_2d_pair(smem + OFF_A0 + 2 * TILE_BYTES, &tmap_n, 0, b * a.L + j0, a0_full_l);
          tma_load_2d_pair(smem + OFF_A0 + 3 * TILE_BYTES, &tmap_n, KBLK, b * a.L + j0, a0_full_l);
        }
        for (int wt = 0; wt < WTILES; ++wt, ++cnt) {
          const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&w_full[s], 2 * HALF_BYTES);
          // the weight image is pre-swizzled: copied verbatim as a [64 rows x 128 B] box of an un-swizzled 2-D view
          tma_load_2d_pair(smem + OFF_W + s * HALF_BYTES, &tmap_w, 0, wrow0 + wt * TM, cluster_addr(&w_full[s], 0));
        }
      }
    }
  } else if (warp == 1) {
    // ===== leader: MMA issuer of the pair (the whole warp runs the loop, one elected lane issues); the peer's warp 1 only
    //       owns its half of the tensor-memory allocation =====
    if (leader) {
      const uint32_t a0 = desc_lo_sw128(smem_u32(smem + OFF_A0)), wr = desc_lo_sw128(smem_u32(smem + OFF_W));
      constexpr uint32_t BLK = TILE_BYTES >> 4;    // activation K-block stride in descriptor units
      constexpr uint32_t HBLK = HALF_BYTES >> 4;   // weight ring slot stride
      uint32_t cnt = 0, ph_a0 = 0, ph_h1 = 0, ph_h2 = 0;
      uint32_t nE = 0, n2 = 0;
      auto next_block = [&]() -> uint32_t {
        const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        return wr + s * HBLK;
      };
      auto release_block = [&]() { umma_commit2(&w_empty[cnt % NSTAGE]); };  // elected lane
      auto wait_prev = [&](uint64_t* bar, uint32_t& n) {  // wait #k waits for drain #(k-1): the first one passes
        mbar_wait(bar, (n & 1) ^ 1);
        ++n;
      };
      for (int pt = pair_id; pt < n_pair_tiles; pt += n_pairs_grid) {
        mbar_wait(a0_full, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        // ---- layer 1: three 128-column chunks in E, R2, E; A = [z | n'_j] from shared memory ----
        for (int nc = 0; nc < 3; ++nc) {
          uint32_t d;
          if (nc == 1) { wait_prev(empty2, n2); d = tmem + 128; }
          else { wait_prev(emptyE, nE); d = tmem; }
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb, ++cnt) {
            const uint32_t wb = next_block();
            if (elect_one()) {
              kblock2_ss(d, a0 + kb * BLK, wb, IDESC256, kb == 0);
              release_block();
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit2(nc == 1 ? full2 : fullE);
          __syncwarp();
        }
        // ---- layer 2: three 128-column chunks in R2, E, R2; A = h1 from tensor memory ----
        for (int c = 0; c < 3; ++c) {
          uint32_t d;
          if (c == 1) { wait_prev(emptyE, nE); d = tmem; }
          else { wait_prev(empty2, n2); d = tmem + 128; }
          tc_fence_after();
          for (int kb = 0; kb < 6; ++kb, ++cnt) {
            if (c == 0 && !(kb & 1)) {
              mbar_wait(&h1p[kb >> 1], ph_h1);
              tc_fence_after();
            }
            const uint32_t wb = next_block();
            if (elect_one()) {
              kblock2_ts(d, tmem + COL_H1 + kb * 32, wb, IDESC256, kb == 0);
              release_block();
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit2(c == 1 ? fullE : full2);
          __syncwarp();
        }
        ph_h1 ^= 1;
        // ---- final layer into F (over the dead h1): [z | n'_j] terms, then h2 ----
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t wb = next_block();
          if (elect_one()) {
            kblock2_ss(tmem + COL_FIN, a0 + kb * BLK, wb, IDESC256, kb == 0);
            release_block();
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit2(a0_empty);  // both CTAs' activation tiles are free
        __syncwarp();
        for (int kb = 0; kb < 6; ++kb, ++cnt) {
          if (!(kb & 1)) {
            mbar_wait(&h2p[kb >> 1], ph_h2);
            tc_fence_after();
          }
          const uint32_t wb = next_block();
          if (elect_one()) {
            kblock2_ts(tmem + COL_FIN, tmem + h2_col(kb), wb, IDESC256, false);
            release_block();
          }
          __syncwarp();
        }
        ph_h2 ^= 1;
        if (elect_one()) umma_commit2(fullF);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue of this CTA's own 128 rows: identical to pair_tc3.cu =====
    const int ew = warp - 2, q = warp & 3, part = ew >> 2, r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t fE = 0, f2 = 0, fF = 0;
    // hand-over barriers live in the leader: shared::cluster addresses (the leader's own window for the leader itself)
    const uint32_t emptyE_l = cluster_addr(emptyE, 0), empty2_l = cluster_addr(empty2, 0);
    const uint32_t h1p_l = cluster_addr(h1p, 0), h2p_l = cluster_addr(h2p, 0);
    // one arrival per warp: every lane has finished its tensor-memory reads / writes (tcgen05.wait + fence) before the
    // __syncwarp that precedes lane 0's release-arrive
    auto hand_over = [&](uint32_t bar_a, uint32_t bar_b) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(bar_a);
        mbar_arrive_cluster(bar_b);
      }
    };
    // The per-residue vectors u_i / p_i and the row's edge mask of tile t + 1 are fetched from global memory while tile t is
    // being processed and parked in the other half of u_s / p_s: between two tiles the epilogue warps then meet on ONE named
    // barrier and start on the (already finished) first accumulator at once.  Before, every tile began with barrier - global
    // loads - barrier, ~1.5 k cycles of exposed latency per tile with the tensor pipe waiting for its accumulators to drain
    // (ncu source page: 13 % of the kernel's stall samples on the index arithmetic and loads of that prologue).
    constexpr int NU = FLAT ? (NSEG * D_ET + 32 * NEW - 1) / (32 * NEW) : 1;
    float un[NU], pn = 0.f, m_next = 0.f;
    auto fetch = [&](int tile) {  // this thread's share of the vectors of `tile`, and its row's mask, into registers
      int bi, jr;
      if constexpr (FLAT) {
        const long f = (long)tile * TM + r;
        bi = (int)(f / a.L);
        jr = (int)(f - (long)bi * a.L);
#pragma unroll
        for (int k = 0; k < NU; ++k) {
          const int c = et + k * 32 * NEW;
          const int sg = c / D_ET;
          un[k] = c < NSEG * D_ET ? a.u[(size_t)(((long)tile * TM + sg * 32) / a.L) * D_ET + (c - sg * D_ET)] : 0.f;
        }
        const int sg = et / C_Z;
        pn = a.p[(size_t)(((long)tile * TM + sg * 32) / a.L) * C_Z + (et - sg * C_Z)];
      } else {
        bi = tile / tiles_per_i;
        jr = (tile % tiles_per_i) * TM + r;
        un[0] = et < D_ET ? a.u[(size_t)bi * D_ET + et] : 0.f;
        pn = et < C_Z ? a.p[(size_t)bi * C_Z + et] : 0.f;
      }
      const int b = bi / a.L;
      m_next = a.mask[bi] * a.mask[(size_t)b * a.L + jr];
    };
    auto stash = [&](uint32_t buf) {  // ... and from the registers into half `buf` of the shared-memory vectors
      float* ud = u_s + buf * NSEG * D_ET;
      float* pd = p_s + buf * NSEG * C_Z;
      if constexpr (FLAT) {
#pragma unroll
        for (int k = 0; k < NU; ++k)
          if (et + k * 32 * NEW < NSEG * D_ET) ud[et + k * 32 * NEW] = un[k];
        pd[et] = pn;
      } else {
        if (et < D_ET) ud[et] = un[0];
        if (et < C_Z) pd[et] = pn;
      }
    };
    if (pair_id < n_pair_tiles) {
      fetch(2 * pair_id + (int)crank);
      stash(0);
    }
    uint32_t it = 0;
    for (int pt = pair_id; pt < n_pair_tiles; pt += n_pairs_grid, ++it) {
      const int tile = 2 * pt + (int)crank;
      // half it & 1 is complete (written during the previous tile or just above) and every warp has left the previous tile,
      // whose half may now be overwritten
      named_bar_sync(1, 32 * NEW);
      const float m = m_next;
      const bool more = pt + n_pairs_grid < n_pair_tiles;
      if (more) fetch(2 * (pt + n_pairs_grid) + (int)crank);  // in flight during layer 1
      const float* u_q = u_s + (it & 1) * NSEG * D_ET + (FLAT ? q * D_ET : 0);
      const float* p_q = p_s + (it & 1) * NSEG * C_Z + (FLAT ? q * C_Z : 0);
      float y[CW];
      auto wait_full = [&](uint64_t* bar, uint32_t& n) {
        mbar_wait(bar, n & 1);
        ++n;
        tc_fence_after();
      };
      auto add_vec = [&](const float* vec) {
#pragma unroll
        for (int e = 0; e < CW; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(vec + e);
          y[e] += t.x; y[e + 1] += t.y; y[e + 2] += t.z; y[e + 3] += t.w;
        }
      };
      auto load_cols = [&](uint32_t col) {
        tmem_ld32_issue(tmem + lane_off + col, y);
        tmem_wait_ld();
      };
      auto store_packed = [&](uint32_t col) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) pk[e] = pack_bf16(fmaxf(y[2 * e], 0.f), fmaxf(y[2 * e + 1], 0.f));
        tmem_st16(tmem + lane_off + col, pk);
      };
      // ---- layer 1: + u_i, relu, pack, into tensor memory as layer 2's A operand ----
      for (int nc = 0; nc < 3; ++nc) {
        const uint32_t base = nc == 1 ? 128u : 0u;
        if (nc == 1) wait_full(full2, f2); else wait_full(fullE, fE);
        load_cols(base + part * CW);
        add_vec(u_q + nc * 128 + part * CW);
        store_packed(COL_H1 + nc * 64 + part * (CW / 2));
        hand_over(nc == 1 ? empty2_l : emptyE_l, h1p_l + nc * 8);
      }
      if (more) stash((it + 1) & 1);
      // ---- layer 2: + b2, relu, pack, in place (chunks 1, 2) or into the spare strip (chunk 0) ----
      for (int c = 0; c < 3; ++c) {
        const uint32_t base = c == 1 ? 0u : 128u;
        if (c == 1) wait_full(fullE, fE); else wait_full(full2, f2);
        load_cols(base + part * CW);
        add_vec(b2_s + c * 128 + part * CW);
        if (c > 0) {
          tc_fence_before();
          named_bar_sync(2 + q, 32 * NPART);
          tc_fence_after();
        }
        store_packed(h2_chunk_col(c) + part * (CW / 2));
        hand_over(c == 1 ? emptyE_l : empty2_l, h2p_l + c * 8);
      }
      // ---- output: + p_i, LayerNorm over 128 channels, * edge mask, bf16 store ----
      {
        wait_full(fullF, fF);
        load_cols(COL_FIN + part * CW);
        tc_fence_before();
        add_vec(p_q + part * CW);
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
  cluster_sync_all();  // no CTA leaves (or frees tensor memory) while the pair's MMAs or remote arrives may still target it
  if (warp == 1) tmem_dealloc2(tmem, 512);
}

}  // namespace

bool edge_transition_pair_supported(int B, int L) { return L % 32 == 0 && (((size_t)B * L * L / TM) % 2 == 0); }

void edge_transition_pair(const EdgeTransitionArgs& a, cudaStream_t st) {
  S2S_CHECK(edge_transition_pair_supported(a.B, a.L), "edge_transition_pair needs L % 32 == 0 and an even number of row tiles");
  static_assert(NEW == 16 && CW == 32, "the in-place h2 stores assume 4 column parts of 32");
  static_assert(32 * NEW == NSEG * C_Z, "p_i staging assumes one element per epilogue thread");
  S2S_CHECK(a.wimg3 && a.nprime_bf16, "edge_transition_pair: weight image / bf16 node embedding missing");
  const bool flat = a.L % TM != 0;
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z_in, rows, C_Z, C_Z);
  const CUtensorMap mn = make_bf16_2d_map(a.nprime_bf16, (size_t)a.B * a.L, C_Z, C_Z, flat ? 32 : 128);
  // the pre-swizzled weight image as [copies * 40 blocks * 128 rows][64] bf16: a 64-row box is one CTA's half of a block, verbatim
  const CUtensorMap mw = make_bf16_2d_map(a.wimg3, (size_t)a.wimg_copies * WTILES * TM, KBLK, KBLK, 64, false);
  Args k;
  k.wimg = a.wimg3; k.u = a.u; k.p = a.p; k.b2 = a.b2; k.ln_w = a.ln_w; k.ln_b = a.ln_b; k.mask = a.mask;
  k.z_out = a.z_out; k.L = a.L; k.n_tiles = (int)(rows / TM); k.ncopy = a.wimg_copies;
  static bool configured = false;
  const int smem = SMEM_BYTES + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  grid &= ~1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ET5_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // persistent kernel: never launch more clusters than can be co-resident (GPCs with an odd SM count strand one SM)
  static int max_clusters = 0;
  if (!max_clusters) {
    S2S_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, edge_transition_pair_kernel<false>, &cfg));
    S2S_CHECK(max_clusters > 0, "edge_transition_pair: no 2-CTA cluster fits");
  }
  if (grid > 2 * max_clusters) grid = 2 * max_clusters;
  cfg.gridDim = dim3(grid);
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see launch_pdl (common.cuh)
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = g_pdl ? 2 : 1;
  if (flat) S2S_CUDA(cudaLaunchKernelEx(&cfg, edge_transition_pair_kernel<true>, mz, mn, mw, k));
  else S2S_CUDA(cudaLaunchKernelEx(&cfg, edge_transition_pair_kernel<false>, mz, mn, mw, k));
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
