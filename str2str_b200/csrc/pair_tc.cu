// Pair-track MLPs on the 5th-gen tensor cores: fused EdgeTransition and fused edge embedder.
//
// Both kernels work on tiles of 128 consecutive pair rows (fixed decoy b and residue i, 128 values of j), keep the
// whole MLP chain on chip and touch HBM once per row for reading and once for writing (bf16):
//   * operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout (TMA tensor maps for the
//     activation tiles, pre-swizzled weight images streamed with 1-D bulk TMA copies),
//   * one elected thread issues tcgen05.mma (M=128, N=128, K=16, bf16 x bf16 -> fp32) with accumulators in TMEM,
//   * four epilogue warps pull accumulators back with tcgen05.ld, apply bias/ReLU/LayerNorm and either re-stage the
//     bf16 activations for the next layer in shared memory or store the output rows.
// Roles talk through mbarriers only.  Requires L % 128 == 0 (the launcher routes other lengths to pair_simt.cu).
//
// EdgeTransition algebra (reference src/models/net/layers.py:170-185), with x = [z_ij, n'_i, n'_j]:
//   h1 = relu(W1 x + b1) = relu(W1[:, z|n'_j] [z_ij, n'_j] + u_i),       u_i = W1[:,128:256] n'_i + b1   (per residue)
//   h2 = relu(W2 h1 + b2)
//   y  = Wf (h2 + x) + bf = [Wf | Wf[:, :128] | Wf[:,256:]] [h2, z_ij, n'_j] + p_i,   p_i = Wf[:,128:256] n'_i + bf
//   out = LayerNorm(y) * mask_i * mask_j
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int ET_RING = 3;
constexpr int ET_WTILES = 40;            // weight blocks per row tile: 12 (layer 1) + 18 (layer 2) + 10 (final)

// ---- fused EdgeTransition -----------------------------------------------------------------------------------------
struct EtTcArgs {
  const bf16* wimg;  // ET_WTILES pre-swizzled [128 x 64] weight blocks in consumption order
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles, ncopy;
};

constexpr int ET_OFF_A0 = 0;                          // [z | n'_j] tile: 4 K-blocks
constexpr int ET_OFF_H = 4 * TILE_BYTES;              // hidden activations: 6 K-blocks
constexpr int ET_OFF_W = ET_OFF_H + 6 * TILE_BYTES;   // weight ring
constexpr int ET_OFF_VEC = ET_OFF_W + ET_RING * TILE_BYTES;
constexpr int ET_VEC_FLOATS = D_ET + C_Z + D_ET + C_Z + C_Z;  // u_i, p_i, b2, ln_w, ln_b
constexpr int ET_OFF_BAR = ET_OFF_VEC + ET_VEC_FLOATS * 4;
constexpr int ET_SMEM = ET_OFF_BAR + 16 * 8 + 16;

__global__ void __launch_bounds__(192, 1)
edge_transition_tc_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n, EtTcArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  float* u_s = reinterpret_cast<float*>(smem + ET_OFF_VEC);
  float* p_s = u_s + D_ET;
  float* b2_s = p_s + C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ET_OFF_BAR);
  uint64_t* w_full = bars;            // [ET_RING]
  uint64_t* w_empty = bars + 3;       // [ET_RING]
  uint64_t* a0_full = bars + 6;
  uint64_t* a0_empty = bars + 7;
  uint64_t* acc_full = bars + 8;
  uint64_t* h_full = bars + 9;
  uint64_t* acc3_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ET_RING; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    mbar_init(acc_full, 1);
    mbar_init(h_full, 128);
    mbar_init(acc3_empty, 128);
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
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_i = a.L / TM;
  constexpr uint32_t IDESC = make_idesc(128, 128);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t cnt = 0, ph_a0 = 0;
      const bf16* wimg = a.wimg + (size_t)(blockIdx.x % a.ncopy) * ((size_t)ET_WTILES * TILE_BYTES / 2);
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
        const int b = bi / a.L;
        mbar_wait(a0_empty, ph_a0 ^ 1);
        mbar_expect_tx(a0_full, 4 * TILE_BYTES);
        tma_load_2d(smem + ET_OFF_A0, &tmap_z, 0, tile * TM, a0_full);
        tma_load_2d(smem + ET_OFF_A0 + TILE_BYTES, &tmap_z, KBLK, tile * TM, a0_full);
        tma_load_2d(smem + ET_OFF_A0 + 2 * TILE_BYTES, &tmap_n, 0, b * a.L + j0, a0_full);
        tma_load_2d(smem + ET_OFF_A0 + 3 * TILE_BYTES, &tmap_n, KBLK, b * a.L + j0, a0_full);
        ph_a0 ^= 1;
        for (int wt = 0; wt < ET_WTILES; ++wt, ++cnt) {
          const uint32_t s = cnt % ET_RING, ph = (cnt / ET_RING) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_expect_tx(&w_full[s], TILE_BYTES);
          tma_bulk_1d(smem + ET_OFF_W + s * TILE_BYTES, wimg + (size_t)wt * (TILE_BYTES / 2), TILE_BYTES, &w_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t a0 = smem_u32(smem + ET_OFF_A0), hb = smem_u32(smem + ET_OFF_H), wr = smem_u32(smem + ET_OFF_W);
      uint32_t cnt = 0, ph_a0 = 0, ph_h = 0, ph_3 = 0;
      auto wblock = [&](uint32_t d_tmem, uint32_t a_blk, bool first) {
        const uint32_t s = cnt % ET_RING, ph = (cnt / ET_RING) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        mma_kblock(d_tmem, a_blk, wr + s * TILE_BYTES, IDESC, first);
        umma_commit(&w_empty[s]);
        ++cnt;
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(a0_full, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        for (int nc = 0; nc < 3; ++nc)  // layer 1: K = [z | n'_j] = 4 blocks
          for (int kb = 0; kb < 4; ++kb) wblock(tmem + nc * 128, a0 + kb * TILE_BYTES, kb == 0);
        umma_commit(acc_full);
        mbar_wait(h_full, ph_h);
        ph_h ^= 1;
        tc_fence_after();
        for (int nc = 0; nc < 3; ++nc)  // layer 2: K = h1 = 6 blocks
          for (int kb = 0; kb < 6; ++kb) wblock(tmem + nc * 128, hb + kb * TILE_BYTES, kb == 0);
        umma_commit(acc_full);
        mbar_wait(h_full, ph_h);
        ph_h ^= 1;
        mbar_wait(acc3_empty, ph_3 ^ 1);  // previous tile's output accumulator has been drained
        ph_3 ^= 1;
        tc_fence_after();
        for (int kb = 0; kb < 6; ++kb) wblock(tmem + 384, hb + kb * TILE_BYTES, kb == 0);  // final: h2
        for (int kb = 0; kb < 4; ++kb) wblock(tmem + 384, a0 + kb * TILE_BYTES, false);    //        z, n'_j
        umma_commit(acc_full);
        umma_commit(a0_empty);
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) =====
    const int q = warp & 3, r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    unsigned char* hbuf = smem + ET_OFF_H;
    uint32_t ph_acc = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / a.L;
      named_bar_sync(1, 128);  // everyone is done with the previous tile's u_i / p_i
      for (int c = et; c < D_ET; c += 128) u_s[c] = a.u[(size_t)bi * D_ET + c];
      p_s[et] = a.p[(size_t)bi * C_Z + et];
      named_bar_sync(1, 128);
      const float m = a.mask[bi] * a.mask[(size_t)b * a.L + j0 + r];
      float v[32];
      // hidden layers: acc -> (+ vec) -> relu -> bf16 -> swizzled operand block for the next MMA
      for (int layer = 0; layer < 2; ++layer) {
        const float* vec = layer == 0 ? u_s : b2_s;
        mbar_wait(acc_full, ph_acc);
        ph_acc ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < D_ET; c0 += 32) {
          tmem_ld32(lane_base + c0, v);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            float h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = fmaxf(v[gq * 8 + e] + vec[c0 + gq * 8 + e], 0.f);
            const int c = c0 + gq * 8;
            store8_sw128(hbuf + (c / KBLK) * TILE_BYTES, r, c % KBLK, h);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(h_full);
      }
      // output layer: + p_i, LayerNorm over 128 channels (exact two-pass), * edge mask, bf16 store
      mbar_wait(acc_full, ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(lane_base + 384 + c0, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) sum += v[e] + p_s[c0 + e];
      }
      const float mean = sum * (1.f / C_Z);
      float sq = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(lane_base + 384 + c0, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = v[e] + p_s[c0 + e] - mean;
          sq += d * d;
        }
      }
      const float rstd = rsqrtf(sq * (1.f / C_Z) + 1e-5f);
      bf16* orow = a.z_out + ((size_t)tile * TM + r) * C_Z;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(lane_base + 384 + c0, v);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c0 + gq * 8 + e;
            o[e] = ((v[gq * 8 + e] + p_s[c] - mean) * rstd * lnw_s[c] + lnb_s[c]) * m;
          }
          *reinterpret_cast<uint4*>(orow + c0 + gq * 8) =
              make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        }
      }
      tc_fence_before();
      mbar_arrive(acc3_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---- fused edge embedder ---------------------------------------------------------------------------------------------
struct EeTcArgs {
  EdgeEmbedArgs e;
  const bf16* wimg;  // 4 pre-swizzled blocks: W2 k0, W2 k1, W3 k0, W3 k1
  int n_tiles;
};
constexpr int WD_PITCH = C_Z + 4;
constexpr int EE_OFF_A = 0;                      // activation tile: 2 K-blocks
constexpr int EE_OFF_W = 2 * TILE_BYTES;         // 4 resident weight blocks
constexpr int EE_OFF_VEC = EE_OFF_W + 4 * TILE_BYTES;
constexpr int EE_VEC_FLOATS = C_Z * 5 + N_BINS * WD_PITCH + 32 + 4 * 128;  // Ti, b2, b3, ln_w, ln_b, Wd, bin edges, LayerNorm partials
constexpr int EE_THREADS = 288;  // MMA/loader warp + 8 worker warps (two per TMEM lane quarter: 16 rows of layer 1, 64 columns of the epilogues each)
constexpr int EE_OFF_BAR = EE_OFF_VEC + EE_VEC_FLOATS * 4;
constexpr int EE_SMEM = EE_OFF_BAR + 8 * 8 + 16;

__global__ void __launch_bounds__(EE_THREADS, 2) edge_embed_tc_kernel(EeTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  float* ti_s = reinterpret_cast<float*>(smem + EE_OFF_VEC);
  float* b2_s = ti_s + C_Z;
  float* b3_s = b2_s + C_Z;
  float* lnw_s = b3_s + C_Z;
  float* lnb_s = lnw_s + C_Z;
  float* wd_s = lnb_s + C_Z;
  float* edge_s = wd_s + N_BINS * WD_PITCH;
  float* red_s = edge_s + 32;  // [2 stats][2 column halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + EE_OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* acc_full = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const EdgeEmbedArgs& e = a.e;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    mbar_init(a_full, 256);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  for (int c = threadIdx.x; c < C_Z; c += blockDim.x) {
    b2_s[c] = e.b2[c];
    b3_s[c] = e.b3[c];
    lnw_s[c] = e.ln_w[c];
    lnb_s[c] = e.ln_b[c];
  }
  for (int c = threadIdx.x; c < N_BINS * C_Z; c += blockDim.x) wd_s[(c / C_Z) * WD_PITCH + (c % C_Z)] = e.Wd[c];
  if (threadIdx.x < N_BINS) edge_s[threadIdx.x] = e.bin_lower[threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_i = e.L / TM;
  constexpr uint32_t IDESC = make_idesc(128, 128);

  if (warp == 0) {
    // loader + MMA issuer: whole warp runs the loop, one elected lane issues
    if (elect_one()) {
      mbar_expect_tx(w_full, 4 * TILE_BYTES);
      for (int t = 0; t < 4; ++t) tma_bulk_1d(smem + EE_OFF_W + t * TILE_BYTES, a.wimg + (size_t)t * (TILE_BYTES / 2), TILE_BYTES, w_full);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    const uint32_t ab = desc_lo_sw128(smem_u32(smem + EE_OFF_A)), wb = desc_lo_sw128(smem_u32(smem + EE_OFF_W));
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    uint32_t ph_a = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(a_full, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        if (elect_one()) {
          kblock_ss(tmem, ab, wb + (2 * layer) * BLK, IDESC, true);
          kblock_ss(tmem, ab + BLK, wb + (2 * layer + 1) * BLK, IDESC, false);
          umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // 8 worker warps: TMEM lane quarter q = warp % 4, hf = which of the quarter's two warps.  (ncu: with 4 workers the
    // kernel was a chain of exposed latencies — layer-1 table reads 45 %, MMA round trips 12 %, TMEM passes 35 % of
    // the samples — so the remedy is more warps in flight and one TMEM pass for the LayerNorm.)
    const int q = warp & 3, hf = (warp - 1) >> 2, r = q * 32 + lane;
    const int wt = threadIdx.x - 32;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    unsigned char* abuf = smem + EE_OFF_A;
    uint32_t ph_acc = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / e.L;
      const size_t bj = (size_t)b * e.L + j0 + r;
      named_bar_sync(1, 256);
      if (wt < C_Z) ti_s[wt] = e.Ti[(size_t)bi * C_Z + wt];
      named_bar_sync(1, 256);
      // layer 1 by table lookups (reference denoising_ipa.py:126-158): time/fixed features of i and j, relative
      // position, self-conditioning distogram bin.  Each lane first classifies one of its quarter's 32 rows, then the
      // warp walks its 16 rows together so that every table row is one coalesced 512-byte read (lane = 4 channels).
      {
        const int bin_l = pair_distogram_bin(e.sc_ca + (size_t)bi * 3, e.sc_ca + bj * 3, edge_s);
        const int off_l = (int)(e.ridx[bi] - e.ridx[bj]) - e.d_min;
        const int c = lane * 4;
        const float4 tiv = *reinterpret_cast<const float4*>(ti_s + c);
        const size_t bj0 = (size_t)b * e.L + j0 + q * 32;
#pragma unroll 8
        for (int r16 = 0; r16 < 16; ++r16) {
          const int rr = hf * 16 + r16;
          const int bin = __shfl_sync(0xffffffffu, bin_l, rr);
          const int off = __shfl_sync(0xffffffffu, off_l, rr);
          const float4 tjv = __ldg(reinterpret_cast<const float4*>(e.Tj + (bj0 + rr) * C_Z + c));
          const float4 tpv = __ldg(reinterpret_cast<const float4*>(e.Tpos + (size_t)off * C_Z + c));
          float4 h = make_float4(tiv.x + tjv.x + tpv.x, tiv.y + tjv.y + tpv.y, tiv.z + tjv.z + tpv.z, tiv.w + tjv.w + tpv.w);
          if (bin >= 0) {
            const float4 wv = *reinterpret_cast<const float4*>(wd_s + bin * WD_PITCH + c);
            h.x += wv.x; h.y += wv.y; h.z += wv.z; h.w += wv.w;
          }
          const int row = q * 32 + rr;
          *reinterpret_cast<uint2*>(abuf + (c / KBLK) * TILE_BYTES + sw128_offset(row, c % KBLK)) =
              make_uint2(pack_bf16(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f)), pack_bf16(fmaxf(h.z, 0.f), fmaxf(h.w, 0.f)));
        }
      }
      fence_proxy_async();
      mbar_arrive(a_full);
      // layer 2 epilogue: + b2, relu, restage (this thread: row r, columns hf*64 .. hf*64+63 = K-block hf)
      float v[64];
      mbar_wait(acc_full, ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      tmem_ld32_issue(lane_base + hf * 64, v);
      tmem_ld32_issue(lane_base + hf * 64 + 32, v + 32);
      tmem_wait_ld();
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        float h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) h[k] = fmaxf(v[gq * 8 + k] + b2_s[hf * 64 + gq * 8 + k], 0.f);
        store8_sw128(abuf + hf * TILE_BYTES, r, gq * 8, h);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(a_full);
      // layer 3 epilogue: + b3, LayerNorm (exact two-pass on registers; the two half-row threads exchange partial
      // sums through shared memory), mask, store
      mbar_wait(acc_full, ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      tmem_ld32_issue(lane_base + hf * 64, v);
      tmem_ld32_issue(lane_base + hf * 64 + 32, v + 32);
      tmem_wait_ld();
      tc_fence_before();  // TMEM reads done before the next tile's MMA (ordered by the next a_full arrive)
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        v[k] += b3_s[hf * 64 + k];
        sum += v[k];
      }
      red_s[hf * 128 + r] = sum;
      named_bar_sync(2 + q, 64);
      const float mean = (sum + red_s[(hf ^ 1) * 128 + r]) * (1.f / C_Z);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        const float d = v[k] - mean;
        sq += d * d;
      }
      red_s[256 + hf * 128 + r] = sq;
      named_bar_sync(2 + q, 64);
      const float rstd = rsqrtf((sq + red_s[256 + (hf ^ 1) * 128 + r]) * (1.f / C_Z) + 1e-5f);
      const float m = e.mask[bi] * e.mask[bj];
      bf16* orow = e.z_out + ((size_t)tile * TM + r) * C_Z + hf * 64;
#pragma unroll
      for (int gq = 0; gq < 8; ++gq) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = hf * 64 + gq * 8 + k;
          o[k] = ((v[gq * 8 + k] - mean) * rstd * lnw_s[c] + lnb_s[c]) * m;
        }
        *reinterpret_cast<uint4*>(orow + gq * 8) =
            make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- host side ------------------------------------------------------------------------------------------------------
__global__ void build_wtile_kernel(const float* __restrict__ src, int ld, int n0, int k0, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TM * KBLK) return;
  const int r = idx / KBLK, c = idx % KBLK;
  *reinterpret_cast<bf16*>(dst + sw128_offset(r, c)) = __float2bfloat16_rn(src[(size_t)(n0 + r) * ld + k0 + c]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    S2S_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    S2S_CHECK(p && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
}  // namespace

// [rows][cols] bf16 row-major tensor (row pitch in elements), boxes of 128 rows x 64 columns, 128-byte swizzle;
// out-of-bounds box elements read as zero
CUtensorMap make_bf16_2d_map(const void* base, size_t rows, size_t cols, size_t row_pitch_elems, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S2S_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
  return m;
}
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    S2S_CUDA(cudaGetDevice(&dev));
    S2S_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

size_t et_wimg_elems() { return (size_t)ET_WTILES * TM * KBLK; }
size_t ee_wimg_elems() { return (size_t)4 * TM * KBLK; }

// Weight images in the order the MMA issuer consumes them (see the kernel's loops).
void build_et_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  auto tile = [&](const float* src, int n0, int k0) {
    build_wtile_kernel<<<TM * KBLK / 256, 256, 0, st>>>(src, D_ET, n0, k0, d);
    S2S_LAUNCH_CHECK();
    d += TILE_BYTES;
  };
  for (int nc = 0; nc < 3; ++nc)
    for (int kb = 0; kb < 4; ++kb) tile(W1, nc * 128, kb < 2 ? kb * KBLK : 256 + (kb - 2) * KBLK);
  for (int nc = 0; nc < 3; ++nc)
    for (int kb = 0; kb < 6; ++kb) tile(W2, nc * 128, kb * KBLK);
  for (int kb = 0; kb < 6; ++kb) tile(Wf, 0, kb * KBLK);
  for (int kb = 0; kb < 2; ++kb) tile(Wf, 0, kb * KBLK);
  for (int kb = 0; kb < 2; ++kb) tile(Wf, 0, 256 + kb * KBLK);
}
void build_ee_wimg(const float* W2, const float* W3, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  const float* srcs[2] = {W2, W3};
  for (int l = 0; l < 2; ++l)
    for (int kb = 0; kb < 2; ++kb) {
      build_wtile_kernel<<<TM * KBLK / 256, 256, 0, st>>>(srcs[l], C_Z, 0, kb * KBLK, d);
      S2S_LAUNCH_CHECK();
      d += TILE_BYTES;
    }
}
void f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t st) {
  f32_to_bf16_kernel<<<ceil_div(n, 256), 256, 0, st>>>(src, dst, n);
  S2S_LAUNCH_CHECK();
}

void edge_transition_tc(const EdgeTransitionArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % TM == 0, "edge_transition_tc needs L % 128 == 0");
  S2S_CHECK(a.wimg && a.nprime_bf16, "edge_transition_tc: weight image / bf16 node embedding missing");
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z_in, rows, C_Z, C_Z);
  const CUtensorMap mn = make_bf16_2d_map(a.nprime_bf16, (size_t)a.B * a.L, C_Z, C_Z);
  EtTcArgs k;
  k.wimg = a.wimg; k.u = a.u; k.p = a.p; k.b2 = a.b2; k.ln_w = a.ln_w; k.ln_b = a.ln_b; k.mask = a.mask;
  k.z_out = a.z_out; k.L = a.L; k.n_tiles = (int)(rows / TM); k.ncopy = a.wimg_copies;
  static bool configured = false;
  const int smem = ET_SMEM + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  const int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  edge_transition_tc_kernel<<<grid, 192, smem, st>>>(mz, mn, k);
  S2S_LAUNCH_CHECK();
}

void edge_embed_tc(const EdgeEmbedArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % TM == 0, "edge_embed_tc needs L % 128 == 0");
  S2S_CHECK(a.wimg, "edge_embed_tc: weight image missing");
  EeTcArgs k;
  k.e = a; k.wimg = a.wimg; k.n_tiles = (int)((size_t)a.B * a.L * a.L / TM);
  static bool configured = false;
  const int smem = EE_SMEM;  // two CTAs per SM: 2 x (112 KB + 1 KB reserved) must stay under 228 KB
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_embed", st);
  const int cap = 2 * sm_count();
  edge_embed_tc_kernel<<<k.n_tiles < cap ? k.n_tiles : cap, EE_THREADS, smem, st>>>(k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
