// Fused EdgeTransition, second generation: MMA and epilogue overlap inside one CTA.
//
// Same math, tile shape (128 pair rows) and operand conventions as pair_tc.cu, re-organised around 64-column chunks:
//   * layers 1 and 2 produce their 384 hidden columns as six 64-wide chunks that ping-pong between two TMEM
//     accumulators, so the four epilogue warps drain chunk c while the tensor pipe computes chunk c+1;
//   * h1 never touches shared memory: the epilogue packs it to bf16 and writes it back to TENSOR MEMORY
//     (tcgen05.st), where layer 2 reads it as the A operand (tcgen05.mma with A in TMEM);
//   * the final layer is accumulated incrementally: as soon as a 64-column chunk of h2 is staged (one 16 KB K-block,
//     double buffered) its partial product with Wf is issued, so h2 is never held whole either;
//   * the [z | n'_j] terms of the final layer are issued right after layer 1, which frees the activation tile early:
//     the next tile's TMA load overlaps layer 2;
//   * shared memory freed by all of the above (224 KB -> 96 KB of activations) becomes a 14-deep ring of 8 KB weight
//     blocks, enough to cover the L2 latency of the streamed weights.
// TMEM map (512 columns): [0,128) two 64-col accumulators | [128,320) h1 as packed bf16 | [320,448) output accumulator.
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int SLOT = 64 * KBLK * 2;  // 8 KiB: one [64 n x 64 k] weight block
constexpr int NSLOT = 14;
constexpr int WTILES = 80;           // 24 (layer 1) + 8 (final: z, n'_j) + 36 (layer 2) + 12 (final: h2)
constexpr int OFF_A0 = 0;
constexpr int OFF_H2 = 4 * TILE_BYTES;
constexpr int OFF_W = OFF_H2 + 2 * TILE_BYTES;
constexpr int OFF_VEC = OFF_W + NSLOT * SLOT;
constexpr int VEC_FLOATS = D_ET + C_Z + D_ET + C_Z + C_Z;  // u_i, p_i, b2, ln_w, ln_b
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
constexpr int N_BARS = 2 * NSLOT + 13;
constexpr int SMEM_BYTES = OFF_BAR + N_BARS * 8 + 16;

constexpr uint32_t COL_ACC = 0, COL_H1 = 128, COL_ACC3 = 320;

struct Args {
  const bf16* wimg;
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles;
};

__global__ void __launch_bounds__(192, 1)
edge_transition_tc2_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n, Args a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  float* u_s = reinterpret_cast<float*>(smem + OFF_VEC);
  float* p_s = u_s + D_ET;
  float* b2_s = p_s + C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NSLOT;
  uint64_t* a0_full = bars + 2 * NSLOT;
  uint64_t* a0_empty = a0_full + 1;
  uint64_t* acc_full = a0_full + 2;    // [2]
  uint64_t* acc_empty = a0_full + 4;   // [2]
  uint64_t* h1_full = a0_full + 6;
  uint64_t* h2_full = a0_full + 7;     // [2]
  uint64_t* h2_empty = a0_full + 9;    // [2]
  uint64_t* acc3_full = a0_full + 11;
  uint64_t* acc3_empty = a0_full + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 128);
      mbar_init(&h2_full[s], 128);
      mbar_init(&h2_empty[s], 1);
    }
    mbar_init(h1_full, 128);
    mbar_init(acc3_full, 1);
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
  constexpr uint32_t IDESC = make_idesc(128, 64);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t cnt = 0, ph_a0 = 0;
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
          const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          mbar_expect_tx(&w_full[s], SLOT);
          tma_bulk_1d(smem + OFF_W + s * SLOT, a.wimg + (size_t)wt * (SLOT / 2), SLOT, &w_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t a0 = smem_u32(smem + OFF_A0), h2b = smem_u32(smem + OFF_H2), wr = smem_u32(smem + OFF_W);
      uint32_t cnt = 0, ph_a0 = 0, ph_h1 = 0, ph_3 = 0, h2cnt = 0;
      // one streamed weight block against an A K-block in shared memory
      auto blk_ss = [&](uint32_t d, uint32_t a_blk, bool first) {
        const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        const uint32_t wb = wr + s * SLOT;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d, smem_desc_sw128(a_blk + k * 32), smem_desc_sw128(wb + k * 32), IDESC, (first && k == 0) ? 0u : 1u);
        umma_commit(&w_empty[s]);
        ++cnt;
      };
      // ... against 64 k-columns of h1 in tensor memory (32 packed columns)
      auto blk_ts = [&](uint32_t d, uint32_t a_col, bool first) {
        const uint32_t s = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        const uint32_t wb = wr + s * SLOT;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(d, a_col + k * 8, smem_desc_sw128(wb + k * 32), IDESC, (first && k == 0) ? 0u : 1u);
        umma_commit(&w_empty[s]);
        ++cnt;
      };
      auto partial = [&]() {  // final-layer partial product of the h2 chunk staged last
        const uint32_t st = h2cnt & 1, ph = (h2cnt >> 1) & 1;
        mbar_wait(&h2_full[st], ph);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) blk_ss(tmem + COL_ACC3 + h * 64, h2b + st * TILE_BYTES, false);
        umma_commit(&h2_empty[st]);
        ++h2cnt;
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(a0_full, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        // layer 1: six chunks, K = [z | n'_j] (4 blocks)
        for (int g = 0; g < 6; ++g) {
          const uint32_t buf = g & 1, ph = (g >> 1) & 1;
          mbar_wait(&acc_empty[buf], ph ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) blk_ss(tmem + COL_ACC + buf * 64, a0 + kb * TILE_BYTES, kb == 0);
          umma_commit(&acc_full[buf]);
        }
        // final layer, [z | n'_j] terms: first writes into the output accumulator
        mbar_wait(acc3_empty, ph_3 ^ 1);
        ph_3 ^= 1;
        tc_fence_after();
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < 4; ++kb) blk_ss(tmem + COL_ACC3 + h * 64, a0 + kb * TILE_BYTES, kb == 0);
        umma_commit(a0_empty);  // activation tile is free: the next tile's TMA overlaps layer 2
        // layer 2: A = h1 in tensor memory; final-layer partials trail one chunk behind
        mbar_wait(h1_full, ph_h1);
        ph_h1 ^= 1;
        tc_fence_after();
        for (int c = 0; c < 6; ++c) {
          const int g = 6 + c;
          const uint32_t buf = g & 1, ph = (g >> 1) & 1;
          mbar_wait(&acc_empty[buf], ph ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < 6; ++kb) blk_ts(tmem + COL_ACC + buf * 64, tmem + COL_H1 + kb * 32, kb == 0);
          umma_commit(&acc_full[buf]);
          if (c > 0) partial();
        }
        partial();
        umma_commit(acc3_full);
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) =====
    const int q = warp & 3, r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t ph_3 = 0, h2cnt = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / a.L;
      named_bar_sync(1, 128);
      for (int c = et; c < D_ET; c += 128) u_s[c] = a.u[(size_t)bi * D_ET + c];
      p_s[et] = a.p[(size_t)bi * C_Z + et];
      named_bar_sync(1, 128);
      const float m = a.mask[bi] * a.mask[(size_t)b * a.L + j0 + r];
      float v[32];
      // layer 1 chunks: + u_i, relu, pack to bf16 pairs, back into tensor memory as layer 2's A operand
      for (int g = 0; g < 6; ++g) {
        const uint32_t buf = g & 1, ph = (g >> 1) & 1;
        mbar_wait(&acc_full[buf], ph);
        tc_fence_after();
        uint32_t pk[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(tmem + lane_off + COL_ACC + buf * 64 + half * 32, v);
          const float* vec = u_s + g * 64 + half * 32;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            pk[half * 16 + e] = pack_bf16(fmaxf(v[2 * e] + vec[2 * e], 0.f), fmaxf(v[2 * e + 1] + vec[2 * e + 1], 0.f));
        }
        tmem_st32(tmem + lane_off + COL_H1 + g * 32, pk);
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
      }
      mbar_arrive(h1_full);
      // layer 2 chunks: + b2, relu, stage as one swizzled K-block for the final-layer partial product
      for (int c = 0; c < 6; ++c) {
        const int g = 6 + c;
        const uint32_t buf = g & 1, ph = (g >> 1) & 1;
        const uint32_t st = h2cnt & 1, ph2 = (h2cnt >> 1) & 1;
        ++h2cnt;
        mbar_wait(&acc_full[buf], ph);
        tc_fence_after();
        mbar_wait(&h2_empty[st], ph2 ^ 1);
        unsigned char* hb = smem + OFF_H2 + st * TILE_BYTES;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(tmem + lane_off + COL_ACC + buf * 64 + half * 32, v);
          const float* vec = b2_s + c * 64 + half * 32;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            float h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = fmaxf(v[gq * 8 + e] + vec[gq * 8 + e], 0.f);
            store8_sw128(hb, r, half * 32 + gq * 8, h);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
        mbar_arrive(&h2_full[st]);
      }
      // output: + p_i, LayerNorm over 128 channels (exact two-pass), * edge mask, bf16 store
      mbar_wait(acc3_full, ph_3);
      ph_3 ^= 1;
      tc_fence_after();
      const uint32_t acc3 = tmem + lane_off + COL_ACC3;
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc3 + c0, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) sum += v[e] + p_s[c0 + e];
      }
      const float mean = sum * (1.f / C_Z);
      float sq = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < C_Z; c0 += 32) {
        tmem_ld32(acc3 + c0, v);
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
        tmem_ld32(acc3 + c0, v);
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

__global__ void build_wtile64_kernel(const float* __restrict__ src, int ld, int n0, int k0, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * KBLK) return;
  const int r = idx / KBLK, c = idx % KBLK;
  *reinterpret_cast<bf16*>(dst + sw128_offset(r, c)) = __float2bfloat16_rn(src[(size_t)(n0 + r) * ld + k0 + c]);
}

}  // namespace

size_t et2_wimg_elems() { return (size_t)WTILES * 64 * KBLK; }

// Weight blocks in exactly the order the MMA issuer consumes them.
void build_et2_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  auto tile = [&](const float* src, int n0, int k0) {
    build_wtile64_kernel<<<64 * KBLK / 256, 256, 0, st>>>(src, D_ET, n0, k0, d);
    S2S_LAUNCH_CHECK();
    d += SLOT;
  };
  auto aug = [](int kb) { return kb < 2 ? kb * KBLK : 256 + (kb - 2) * KBLK; };  // [z | n'_j] columns of a 384-wide weight
  for (int g = 0; g < 6; ++g)
    for (int kb = 0; kb < 4; ++kb) tile(W1, g * 64, aug(kb));
  for (int h = 0; h < 2; ++h)
    for (int kb = 0; kb < 4; ++kb) tile(Wf, h * 64, aug(kb));
  for (int c = 0; c < 6; ++c) {
    for (int kb = 0; kb < 6; ++kb) tile(W2, c * 64, kb * KBLK);
    if (c > 0)
      for (int h = 0; h < 2; ++h) tile(Wf, h * 64, (c - 1) * KBLK);
  }
  for (int h = 0; h < 2; ++h) tile(Wf, h * 64, 5 * KBLK);
}

void edge_transition_tc2(const EdgeTransitionArgs& a, cudaStream_t st) {
  S2S_CHECK(a.L % TM == 0, "edge_transition_tc2 needs L % 128 == 0");
  S2S_CHECK(a.wimg2 && a.nprime_bf16, "edge_transition_tc2: weight image / bf16 node embedding missing");
  const size_t rows = (size_t)a.B * a.L * a.L;
  const CUtensorMap mz = make_bf16_2d_map(a.z_in, rows, C_Z, C_Z);
  const CUtensorMap mn = make_bf16_2d_map(a.nprime_bf16, (size_t)a.B * a.L, C_Z, C_Z);
  Args k;
  k.wimg = a.wimg2; k.u = a.u; k.p = a.p; k.b2 = a.b2; k.ln_w = a.ln_w; k.ln_b = a.ln_b; k.mask = a.mask;
  k.z_out = a.z_out; k.L = a.L; k.n_tiles = (int)(rows / TM);
  static bool configured = false;
  const int smem = SMEM_BYTES + 1024;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  const int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  edge_transition_tc2_kernel<<<grid, 192, smem, st>>>(mz, mn, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
