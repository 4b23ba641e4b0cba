// Second-generation fused edge embedder (reference src/models/net/denoising_ipa.py:126-158, geo_utils.py:44-56).
//
// The first-generation kernel (pair_tc.cu: edge_embed_tc_kernel) walks every 128-row tile through its five steps in
// lock-step (table gather -> MMA -> bias/ReLU restage -> MMA -> LayerNorm/store), so each step's latency is exposed and the
// kernel sits at ~17 % of either roofline (1.16 ms at B = 64, L = 256: 1.07 GB written, 0.27 TFLOP).  Here the steps are
// stations of a pipeline, each owned by its own warps and each working on a DIFFERENT tile at any moment:
//
//   G  (8 warps)  layer 1 by table lookups for tile t+2  -> A1[stage]  (bf16, K-major SWIZZLE_128B, 2 stages)
//   M  (1 warp)   tcgen05.mma  layer 2 of tile t+1 (A1 x W2 -> acc2[stage]), layer 3 of tile t (A2 x W3 -> acc3[stage])
//   E2 (4 warps)  acc2 -> + b2, ReLU -> A2[stage]                                (2 stages)
//   E3 (4 warps)  acc3 -> + b3, LayerNorm (two-pass, re-reading tensor memory), mask -> z (bf16) in HBM
//
// One persistent CTA per SM (tiles strided by the grid, so the SMs write neighbouring tiles at any moment); the four
// accumulators fill the 512 columns of tensor memory; W2 / W3 stay resident in shared memory; stations hand tiles over
// through mbarriers only.  Same arithmetic and rounding points as the first generation (and as pair_simt.cu).
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int P_G_WARPS = 8;
constexpr int P_THREADS = 32 * (1 + P_G_WARPS + 4 + 4);
constexpr int P_WD_PITCH = C_Z + 4;
constexpr int P_OFF_A1 = 0;                    // 2 stages x 2 K-blocks
constexpr int P_OFF_A2 = 4 * TILE_BYTES;       // 2 stages x 2 K-blocks
constexpr int P_OFF_W = 8 * TILE_BYTES;        // W2 k0, W2 k1, W3 k0, W3 k1
constexpr int P_OFF_VEC = 12 * TILE_BYTES;
constexpr int P_VEC_FLOATS = 4 * C_Z + N_BINS * P_WD_PITCH + 32;  // b2, b3, ln_w, ln_b, Wd, bin edges
constexpr int P_OFF_BAR = P_OFF_VEC + P_VEC_FLOATS * 4;
constexpr int P_SMEM = P_OFF_BAR + 20 * 8 + 16;

struct EePipeArgs {
  EdgeEmbedArgs e;
  const bf16* wimg;
  int n_tiles;
};

__global__ void __launch_bounds__(P_THREADS, 1) edge_embed_pipe_kernel(EePipeArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* b2_s = reinterpret_cast<float*>(smem + P_OFF_VEC);
  float* b3_s = b2_s + C_Z;
  float* lnw_s = b3_s + C_Z;
  float* lnb_s = lnw_s + C_Z;
  float* wd_s = lnb_s + C_Z;
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
      mbar_init(&a2_full[s], 128);
      mbar_init(&a2_empty[s], 1);
      mbar_init(&acc3_full[s], 1);
      mbar_init(&acc3_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  for (int c = threadIdx.x; c < C_Z; c += blockDim.x) {
    b2_s[c] = e.b2[c];
    b3_s[c] = e.b3[c];
    lnw_s[c] = e.ln_w[c];
    lnb_s[c] = e.ln_b[c];
  }
  for (int c = threadIdx.x; c < N_BINS * C_Z; c += blockDim.x) wd_s[(c / C_Z) * P_WD_PITCH + (c % C_Z)] = e.Wd[c];
  if (threadIdx.x < N_BINS) edge_s[threadIdx.x] = e.bin_lower[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_i = e.L / TM;
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
        mbar_wait(&a1_full[s], ph);
        mbar_wait(&acc2_empty[s], ph ^ 1);
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
        mbar_wait(&a2_full[s], ph);
        mbar_wait(&acc3_empty[s], ph ^ 1);
        tc_fence_after();
        if (elect_one()) {
          kblock_ss(tmem + 256 + s * 128, a2 + s * 2 * BLK, wb + 2 * BLK, IDESC, true);
          kblock_ss(tmem + 256 + s * 128, a2 + s * 2 * BLK + BLK, wb + 3 * BLK, IDESC, false);
          umma_commit(&a2_empty[s]);
          umma_commit(&acc3_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp <= P_G_WARPS) {
    // ---- G: layer 1 by table lookups.  Warp g owns rows g*16 .. g*16+15 of the tile; lanes first classify those rows
    // (distogram bin, relative-position offset), then the warp walks the rows together so that every table row is one
    // coalesced 512-byte read (lane = 4 channels).
    const int g = warp - 1;
    const int c = lane * 4;
    unsigned char* const dst0 = smem + P_OFF_A1 + (c / KBLK) * TILE_BYTES;
    for (int k = 0; k < n_local; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / e.L;
      const size_t bj0 = (size_t)b * e.L + j0 + g * 16;
      const size_t bjl = bj0 + (lane & 15);
      const int bin_l = pair_distogram_bin(e.sc_ca + (size_t)bi * 3, e.sc_ca + bjl * 3, edge_s);
      const int off_l = (int)(e.ridx[bi] - e.ridx[bjl]) - e.d_min;
      const float4 tiv = __ldg(reinterpret_cast<const float4*>(e.Ti + (size_t)bi * C_Z + c));
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      mbar_wait(&a1_empty[s], ph ^ 1);
      unsigned char* const dst = dst0 + s * 2 * TILE_BYTES;
#pragma unroll 8
      for (int r16 = 0; r16 < 16; ++r16) {
        const int bin = __shfl_sync(0xffffffffu, bin_l, r16);
        const int off = __shfl_sync(0xffffffffu, off_l, r16);
        const float4 tjv = __ldg(reinterpret_cast<const float4*>(e.Tj + (bj0 + r16) * C_Z + c));
        const float4 tpv = __ldg(reinterpret_cast<const float4*>(e.Tpos + (size_t)off * C_Z + c));
        float4 h = make_float4(tiv.x + tjv.x + tpv.x, tiv.y + tjv.y + tpv.y, tiv.z + tjv.z + tpv.z, tiv.w + tjv.w + tpv.w);
        if (bin >= 0) {
          const float4 wv = *reinterpret_cast<const float4*>(wd_s + bin * P_WD_PITCH + c);
          h.x += wv.x; h.y += wv.y; h.z += wv.z; h.w += wv.w;
        }
        *reinterpret_cast<uint2*>(dst + sw128_offset(g * 16 + r16, c % KBLK)) =
            make_uint2(pack_bf16(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f)), pack_bf16(fmaxf(h.z, 0.f), fmaxf(h.w, 0.f)));
      }
      fence_proxy_async();
      mbar_arrive(&a1_full[s]);
    }
  } else if (warp <= P_G_WARPS + 4) {
    // ---- E2: acc2 -> + b2, ReLU -> A2 (this thread: row r of the tile, all 128 columns in two halves) ----
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    for (int k = 0; k < n_local; ++k) {
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      mbar_wait(&acc2_full[s], ph);
      mbar_wait(&a2_empty[s], ph ^ 1);
      tc_fence_after();
      unsigned char* const abuf = smem + P_OFF_A2 + s * 2 * TILE_BYTES;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float v[64];
        tmem_ld32_issue(lane_base + s * 128 + hf * 64, v);
        tmem_ld32_issue(lane_base + s * 128 + hf * 64 + 32, v + 32);
        tmem_wait_ld();
        if (hf == 1) {  // both halves are in registers: the accumulator can take the next tile
          tc_fence_before();
          mbar_arrive(&acc2_empty[s]);
        }
#pragma unroll
        for (int gq = 0; gq < 8; ++gq) {
          float h[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) h[u] = fmaxf(v[gq * 8 + u] + b2_s[hf * 64 + gq * 8 + u], 0.f);
          store8_sw128(abuf + hf * TILE_BYTES, r, gq * 8, h);
        }
      }
      fence_proxy_async();
      mbar_arrive(&a2_full[s]);
    }
  } else {
    // ---- E3: acc3 -> + b3, LayerNorm (exact two-pass; tensor memory is re-read instead of holding 128 values), mask, store
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t lane_base = tmem + 256 + ((uint32_t)(q * 32) << 16);
    for (int k = 0; k < n_local; ++k) {
      const int tile = blockIdx.x + k * gridDim.x;
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const size_t bj = (size_t)(bi / e.L) * e.L + j0 + r;
      const float m = __ldg(e.mask + bi) * __ldg(e.mask + bj);
      const uint32_t s = k & 1, ph = (k >> 1) & 1;
      mbar_wait(&acc3_full[s], ph);
      tc_fence_after();
      const uint32_t acc = lane_base + s * 128;
      float v[32];
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc + c0, v);
#pragma unroll
        for (int u = 0; u < 32; ++u) sum += v[u] + b3_s[c0 + u];
      }
      const float mean = sum * (1.f / C_Z);
      float sq = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc + c0, v);
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const float d = v[u] + b3_s[c0 + u] - mean;
          sq += d * d;
        }
      }
      const float rstd = rsqrtf(sq * (1.f / C_Z) + 1e-5f);
      bf16* orow = e.z_out + ((size_t)tile * TM + r) * C_Z;
#pragma unroll 1
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
            o[u] = ((v[gq * 8 + u] + b3_s[cc] - mean) * rstd * lnw_s[cc] + lnb_s[cc]) * m;
          }
          *reinterpret_cast<uint4*>(orow + c0 + gq * 8) =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

void edge_embed_tc2(const EdgeEmbedArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % TM == 0, "edge_embed_tc2 needs L % 128 == 0");
  S2S_CHECK(a.wimg, "edge_embed_tc2: weight image missing");
  EePipeArgs k;
  k.e = a; k.wimg = a.wimg; k.n_tiles = (int)((size_t)a.B * a.L * a.L / TM);
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
    configured = true;
  }
  S2S_PROF("edge_embed", st);
  const int cap = sm_count();
  edge_embed_pipe_kernel<<<k.n_tiles < cap ? k.n_tiles : cap, P_THREADS, P_SMEM, st>>>(k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
