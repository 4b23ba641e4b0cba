// Fused EdgeTransition, second generation: MMA and epilogue overlap inside one CTA.
//
// Same math, tile shape (128 pair rows) and operand conventions as pair_tc.cu.  What changed, and why (measured on B200,
// see profiles/README.md): the first-generation kernel ran its three layers strictly serially (MMA, then epilogue, then
// MMA ...) and spent ~70 % of each tile outside the tensor pipe.  Here
//   * every layer produces its columns as 128-wide chunks that ping-pong between two TMEM accumulators, so the four
//     epilogue warps drain chunk c while the tensor pipe computes chunk c+1 (and the LayerNorm/store epilogue of a tile
//     overlaps layer 1 of the next tile);
//   * h1 never touches shared memory: the epilogue packs it to bf16 and writes it back to TENSOR MEMORY (tcgen05.st),
//     where layer 2 reads it as the A operand (tcgen05.mma with A in TMEM) — this halves the shared-memory read traffic
//     of the largest layer, which at N=128 is otherwise right at the 128 B/clk shared-memory limit;
//   * weights still stream as pre-swizzled 16 KB blocks [128 n x 64 k]: an ablation (S2S_ET_DEBUG) showed the stream is
//     not a bottleneck, but that per-block barrier traffic of the single MMA-issuing thread is, so blocks are kept as
//     large as the operand layout allows.
// TMEM map (512 columns): [0,256) two 128-col fp32 accumulators | [256,448) h1 as packed bf16 (A operand of layer 2).
#include <cstdlib>

#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int NSTAGE = 3;
constexpr int WTILES = 40;  // 12 (layer 1) + 18 (layer 2) + 10 (final)
constexpr int OFF_A0 = 0;                       // [z | n'_j] tile: 4 K-blocks
constexpr int OFF_H2 = 4 * TILE_BYTES;          // h2: 6 K-blocks
constexpr int OFF_W = OFF_H2 + 6 * TILE_BYTES;  // weight ring
constexpr int OFF_VEC = OFF_W + NSTAGE * TILE_BYTES;
constexpr int VEC_FLOATS = D_ET + C_Z + D_ET + C_Z + C_Z + 512;  // u_i, p_i, b2, ln_w, ln_b, LayerNorm partial sums
constexpr int OFF_BAR = OFF_VEC + VEC_FLOATS * 4;
constexpr int N_BARS = 2 * NSTAGE + 9;
constexpr int SMEM_BYTES = OFF_BAR + N_BARS * 8 + 16;

constexpr uint32_t COL_ACC = 0, COL_H1 = 256;

struct Args {
  const bf16* wimg;
  const float *u, *p, *b2, *ln_w, *ln_b, *mask;
  bf16* z_out;
  int L, n_tiles, ncopy;
  int dbg;  // timing experiments only (S2S_ET_DEBUG): 1 no weight TMA, 2 no MMA, 4 no epilogue math
};

__global__ void __launch_bounds__(320, 1)
edge_transition_tc2_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_n, Args a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  float* u_s = reinterpret_cast<float*>(smem + OFF_VEC);
  float* p_s = u_s + D_ET;
  float* b2_s = p_s + C_Z;
  float* lnw_s = b2_s + D_ET;
  float* lnb_s = lnw_s + C_Z;
  float* red_s = lnb_s + C_Z;  // [2 stats][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;
  uint64_t* w_empty = bars + NSTAGE;
  uint64_t* a0_full = bars + 2 * NSTAGE;
  uint64_t* a0_empty = a0_full + 1;
  uint64_t* acc_full = a0_full + 2;   // [2]
  uint64_t* acc_empty = a0_full + 4;  // [2]
  uint64_t* h1_full = a0_full + 6;
  uint64_t* h2_full = a0_full + 7;
  uint64_t* h2_empty = a0_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 256);
    }
    mbar_init(h1_full, 256);
    mbar_init(h2_full, 256);
    mbar_init(h2_empty, 1);
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
          tma_bulk_1d(smem + OFF_W + s * TILE_BYTES, wimg + (size_t)wt * (TILE_BYTES / 2), TILE_BYTES, &w_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t a0 = smem_u32(smem + OFF_A0), h2b = smem_u32(smem + OFF_H2), wr = smem_u32(smem + OFF_W);
      uint32_t cnt = 0, ph_a0 = 0, ph_h1 = 0, ph_h2 = 0, chunk = 0;
      const bool do_mma = !(a.dbg & 2);
      // one streamed weight block (64 k) against an A K-block in shared memory / 32 packed h1 columns in tensor memory
      auto blk = [&](uint32_t d, uint32_t a_src, bool a_in_tmem, bool first) {
        const uint32_t s = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1;
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        const uint32_t wb = wr + s * TILE_BYTES;
        if (do_mma) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (first && k == 0) ? 0u : 1u;
            if (a_in_tmem) umma_bf16_ts(d, a_src + k * 8, smem_desc_sw128(wb + k * 32), IDESC, acc);
            else umma_bf16(d, smem_desc_sw128(a_src + k * 32), smem_desc_sw128(wb + k * 32), IDESC, acc);
          }
        }
        umma_commit(&w_empty[s]);
        ++cnt;
      };
      // claim the next accumulator of the ping-pong pair
      auto claim = [&]() -> uint32_t {
        const uint32_t buf = chunk & 1, ph = (chunk >> 1) & 1;
        mbar_wait(&acc_empty[buf], ph ^ 1);
        tc_fence_after();
        return buf;
      };
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        mbar_wait(a0_full, ph_a0);
        ph_a0 ^= 1;
        tc_fence_after();
        for (int nc = 0; nc < 3; ++nc, ++chunk) {  // layer 1: K = [z | n'_j] (4 blocks), A in shared memory
          const uint32_t buf = claim();
          for (int kb = 0; kb < 4; ++kb) blk(tmem + COL_ACC + buf * 128, a0 + kb * TILE_BYTES, false, kb == 0);
          umma_commit(&acc_full[buf]);
        }
        mbar_wait(h1_full, ph_h1);
        ph_h1 ^= 1;
        mbar_wait(h2_empty, ph_h2 ^ 1);  // previous tile's final layer has consumed h2
        tc_fence_after();
        for (int nc = 0; nc < 3; ++nc, ++chunk) {  // layer 2: K = h1 (6 blocks), A in tensor memory
          const uint32_t buf = claim();
          for (int kb = 0; kb < 6; ++kb) blk(tmem + COL_ACC + buf * 128, tmem + COL_H1 + kb * 32, true, kb == 0);
          umma_commit(&acc_full[buf]);
        }
        {  // final layer: [z | n'_j] terms first (frees the activation tile for the next TMA), then h2
          const uint32_t buf = claim();
          for (int kb = 0; kb < 4; ++kb) blk(tmem + COL_ACC + buf * 128, a0 + kb * TILE_BYTES, false, kb == 0);
          umma_commit(a0_empty);
          mbar_wait(h2_full, ph_h2);
          ph_h2 ^= 1;
          tc_fence_after();
          for (int kb = 0; kb < 6; ++kb) blk(tmem + COL_ACC + buf * 128, h2b + kb * TILE_BYTES, false, false);
          umma_commit(&acc_full[buf]);
          umma_commit(h2_empty);
          ++chunk;
        }
      }
    }
  } else {
    // ===== 8 epilogue warps: TMEM lane quarter = warp % 4, two warps per quarter split each chunk's 128 columns =====
    const int ew = warp - 2, q = warp & 3, hf = ew >> 2, r = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool do_epi = !(a.dbg & 4);
    uint32_t chunk = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int bi = tile / tiles_per_i, j0 = (tile % tiles_per_i) * TM;
      const int b = bi / a.L;
      named_bar_sync(1, 256);
      for (int c = et; c < D_ET; c += 256) u_s[c] = a.u[(size_t)bi * D_ET + c];
      if (et < C_Z) p_s[et] = a.p[(size_t)bi * C_Z + et];
      named_bar_sync(1, 256);
      const float m = a.mask[bi] * a.mask[(size_t)b * a.L + j0 + r];
      float y[64];  // this thread's half (64 columns) of one accumulator chunk of its row
      auto fetch = [&]() -> uint32_t {  // wait for the next chunk and pull it into registers
        const uint32_t buf = chunk & 1, ph = (chunk >> 1) & 1;
        ++chunk;
        mbar_wait(&acc_full[buf], ph);
        tc_fence_after();
        if (do_epi) {
          tmem_ld32_issue(tmem + lane_off + COL_ACC + buf * 128 + hf * 64, y);
          tmem_ld32_issue(tmem + lane_off + COL_ACC + buf * 128 + hf * 64 + 32, y + 32);
          tmem_wait_ld();
        }
        return buf;
      };
      auto add_vec = [&](const float* vec) {  // y += vec[0..64) (broadcast shared-memory reads, 128-bit)
#pragma unroll
        for (int e = 0; e < 64; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(vec + e);
          y[e] += t.x; y[e + 1] += t.y; y[e + 2] += t.z; y[e + 3] += t.w;
        }
      };
      // layer 1: + u_i, relu, pack to bf16 pairs, back into tensor memory as layer 2's A operand
      for (int nc = 0; nc < 3; ++nc) {
        const uint32_t buf = fetch();
        if (do_epi) {
          add_vec(u_s + nc * 128 + hf * 64);
          uint32_t pk[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) pk[e] = pack_bf16(fmaxf(y[2 * e], 0.f), fmaxf(y[2 * e + 1], 0.f));
          tmem_st32(tmem + lane_off + COL_H1 + nc * 64 + hf * 32, pk);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
      }
      mbar_arrive(h1_full);
      // layer 2: + b2, relu, stage h2 as swizzled K-blocks for the final layer
      for (int nc = 0; nc < 3; ++nc) {
        const uint32_t buf = fetch();
        if (do_epi) {
          add_vec(b2_s + nc * 128 + hf * 64);
          unsigned char* kblk = smem + OFF_H2 + (nc * 2 + hf) * TILE_BYTES;  // this half is exactly one 64-wide K-block
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            float h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = fmaxf(y[gq * 8 + e], 0.f);
            store8_sw128(kblk, r, gq * 8, h);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
      }
      mbar_arrive(h2_full);
      // output: + p_i, LayerNorm over 128 channels (exact two-pass; the two half-row threads exchange partial sums
      // through shared memory), * edge mask, bf16 store
      {
        const uint32_t buf = fetch();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);  // accumulator is in registers: the tensor pipe may reuse it
        if (!do_epi) continue;
        add_vec(p_s + hf * 64);
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 64; ++e) sum += y[e];
        red_s[hf * 128 + r] = sum;
        named_bar_sync(2 + q, 64);
        const float mean = (sum + red_s[(hf ^ 1) * 128 + r]) * (1.f / C_Z);
        float sq = 0.f;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          const float d = y[e] - mean;
          sq += d * d;
        }
        red_s[256 + hf * 128 + r] = sq;
        named_bar_sync(2 + q, 64);
        const float rstd = rsqrtf((sq + red_s[256 + (hf ^ 1) * 128 + r]) * (1.f / C_Z) + 1e-5f);
        bf16* orow = a.z_out + ((size_t)tile * TM + r) * C_Z + hf * 64;
        const float* lw = lnw_s + hf * 64;
        const float* lb = lnb_s + hf * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 8) {
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
  if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void build_wtile128_kernel(const float* __restrict__ src, int ld, int n0, int k0, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TM * KBLK) return;
  const int r = idx / KBLK, c = idx % KBLK;
  *reinterpret_cast<bf16*>(dst + sw128_offset(r, c)) = __float2bfloat16_rn(src[(size_t)(n0 + r) * ld + k0 + c]);
}

}  // namespace

size_t et2_wimg_elems() { return (size_t)WTILES * TM * KBLK; }

// Weight blocks in exactly the order the MMA issuer consumes them.
void build_et2_wimg(const float* W1, const float* W2, const float* Wf, bf16* dst, cudaStream_t st) {
  unsigned char* d = reinterpret_cast<unsigned char*>(dst);
  auto tile = [&](const float* src, int n0, int k0) {
    build_wtile128_kernel<<<TM * KBLK / 256, 256, 0, st>>>(src, D_ET, n0, k0, d);
    S2S_LAUNCH_CHECK();
    d += TILE_BYTES;
  };
  auto aug = [](int kb) { return kb < 2 ? kb * KBLK : 256 + (kb - 2) * KBLK; };  // [z | n'_j] columns of a 384-wide weight
  for (int nc = 0; nc < 3; ++nc)
    for (int kb = 0; kb < 4; ++kb) tile(W1, nc * 128, aug(kb));
  for (int nc = 0; nc < 3; ++nc)
    for (int kb = 0; kb < 6; ++kb) tile(W2, nc * 128, kb * KBLK);
  for (int kb = 0; kb < 4; ++kb) tile(Wf, 0, aug(kb));
  for (int kb = 0; kb < 6; ++kb) tile(Wf, 0, kb * KBLK);
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
    S2S_CUDA(cudaFuncSetAttribute(edge_transition_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  S2S_PROF("edge_transition", st);
  const int grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  edge_transition_tc2_kernel<<<grid, 320, smem, st>>>(mz, mn, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
