// Fused chains of node-track linear layers:  x_{s+1} = epilogue_s(x_s W_s^T)  for up to CHAIN_MAX_STEPS consecutive layers
// that act row by row (reference ipa.py:353-365 — the sequence transformer's out_proj / norm1 / linear1 / linear2 / norm2,
// the following in_proj or the post-transformer linear, layers.py:138-145 NodeTransition, layers.py:170-176 the per-residue
// part of EdgeTransition, layers.py:199-213 the torsion head), in ONE launch.
//
// Why: at cfg2 every one of these GEMMs is a 16384 x (256..960) x (256..320) problem that takes 17-21 us as its own launch,
// of which ~14 us is fixed cost (launch, barrier / tensor-memory prologue, pipeline fill, last epilogue) and per-CTA weight
// streaming (profiles/r01c_gemm_and_embedder_experiments.log, section 3), and each LayerNorm between them is one more launch
// that re-reads and re-writes the activations.  All of them are ROW-LOCAL: a CTA that owns a 128-row panel can run the whole
// chain on it without ever looking at another CTA's rows, so no grid-wide synchronisation is needed between the layers.
//
// One step is the panel GEMM of gemm_tc.cu (A resident in tensor memory as split-bf16 hi | lo images, TS-mode MMAs, weight
// blocks streamed through a 12-block shared-memory ring, two accumulator buffers, eight epilogue warps with the coalesced
// compile-time-specialised epilogue), always as the 3-pass split-bf16 product.  What is new:
//   * the step loop: after the last chunk of a step the epilogue threads make their global stores visible to the async proxy
//     (fence.proxy.async) and arrive on `step_done`; the TMA producer waits for it before it fetches the next step's A panel —
//     which is what this CTA has just written (L2-resident) — so a step boundary costs one panel reload instead of a kernel
//     boundary.  (Measured alternative, reverted: the eight epilogue warps reading the panel back with plain loads so that the
//     weight ring could run ahead across steps — row-per-thread 16-byte loads expose more latency than the TMA boxes: 114.7 vs
//     122.1 conformations/s at L = 64 x 32 decoys, 48.96 vs 49.44 at cfg2.)
//   * LayerNorm as a step epilogue: the eight epilogue warps meet on a named barrier once every chunk of the step is stored,
//     then each normalises 16 of the panel's rows (layernorm_rows of row_ops.cuh — the same code, hence the same bits, as the
//     stand-alone LayerNorm kernel) and writes the fp32 row and its split-bf16 image for the next step.
// sm_100a only.
#include <cstring>
#include <type_traits>

#include "row_ops.cuh"
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int C_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps (which also stage A into tensor memory)
constexpr int C_RING = 12 * TILE_BYTES;
constexpr int C_OFF_BAR = C_RING;
constexpr int C_OFF_BIAS = C_OFF_BAR + 32 * 8 + 16;  // per-epilogue-warp bias slice, [8][64] fp32
constexpr int C_OFF_XP = C_OFF_BIAS + 8 * 64 * 4;    // per-epilogue-warp 32 x 32 fp32 transposition buffer
constexpr int C_SMEM = C_OFF_XP + 8 * 32 * 32 * 4;
constexpr int C_STAGES = 6;                          // ring stages of (hi, lo) block pairs

struct ChainK {  // one step as the kernel sees it
  int N, K, relu, epi;           // epi: bit 0 fp32 C, 1 bf16 hi image, 2 lo image, 3 residual added, 4 bf16-only wide path
  int acol_lo, acc0, acc1, w1;   // tensor-memory geometry of this step (A_hi at column 0)
  const float *bias, *res, *row_post;
  float* C;
  bf16 *out_hi, *out_lo;
  long ldc, ldres, ldo;
  const float *ln_w, *ln_b, *ln_scale;  // LayerNorm over the N columns of C (when ln_w is set)
  float* ln_out;
  bf16 *ln_hi, *ln_lo;
  long ld_ln;
};
struct ChainKArgs {
  int M, n_steps;
  ChainK s[CHAIN_MAX_STEPS];
};
struct ChainMaps {
  CUtensorMap m[CHAIN_MAX_STEPS][4];  // A_hi, A_lo, W_hi, W_lo
};

__global__ void __launch_bounds__(C_THREADS, 1)
gemm_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainKArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_OFF_BAR);
  uint64_t* s_full = bars;          // [6]
  uint64_t* s_empty = bars + 12;    // [6]
  uint64_t* acc_full = bars + 24;   // [2]
  uint64_t* acc_empty = bars + 26;  // [2]
  uint64_t* a_ready = bars + 28;    // A of the current step is in tensor memory
  uint64_t* a_free = bars + 29;     // every MMA of the step has completed: A may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  uint64_t* step_done = bars + 31;  // every output of the step is in global memory, visible to the async proxy

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  constexpr uint32_t stage_bytes = 2 * TILE_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 32 * 8);
    }
    mbar_init(a_ready, 32 * 8);
    mbar_init(a_free, 1);
    mbar_init(step_done, 32 * 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();
  const int MT = (a.M + TM - 1) / TM;
  // chunk ci of the whole CTA run uses accumulator buffer ci & 1; its width follows from the buffer and what is left of N
  auto chunk_width = [](const ChainK& s, uint32_t ci, int n_done) { return min((ci & 1) ? s.w1 : 128, s.N - n_done); };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0, ci = 0, g = 0;
      for (int p = blockIdx.x; p < MT; p += gridDim.x) {
        for (int si = 0; si < a.n_steps; ++si, ++g) {
          const ChainK& s = a.s[si];
          const int KB = s.K / KBLK;
          // the A panel of every step but the first is what this CTA's epilogue wrote during the previous step
          if (g > 0) mbar_wait(step_done, (g - 1) & 1);
          for (int kb = 0; kb < KB; ++kb, ++cnt) {
            const uint32_t st_i = cnt % C_STAGES, ph = (cnt / C_STAGES) & 1;
            mbar_wait(&s_empty[st_i], ph ^ 1);
            mbar_expect_tx(&s_full[st_i], stage_bytes);
            unsigned char* st = smem + st_i * stage_bytes;
            tma_load_2d(st, &maps.m[si][0], kb * KBLK, p * TM, &s_full[st_i]);
            tma_load_2d(st + TILE_BYTES, &maps.m[si][1], kb * KBLK, p * TM, &s_full[st_i]);
          }
          for (int n = 0; n < s.N; ++ci) {  // weight blocks: 128 rows of W from row n (rows past the end read as zeros), one K block
            const int w = chunk_width(s, ci, n);
            for (int kb = 0; kb < KB; ++kb, ++cnt) {
              const uint32_t st_i = cnt % C_STAGES, ph = (cnt / C_STAGES) & 1;
              mbar_wait(&s_empty[st_i], ph ^ 1);
              mbar_expect_tx(&s_full[st_i], stage_bytes);
              unsigned char* st = smem + st_i * stage_bytes;
              tma_load_2d(st, &maps.m[si][2], kb * KBLK, n, &s_full[st_i]);
              tma_load_2d(st + TILE_BYTES, &maps.m[si][3], kb * KBLK, n, &s_full[st_i]);
            }
            n += w;
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (whole warp converged, one elected lane issues): D[acc] += A[tmem] * W_blk^T, three split-bf16 passes
    uint32_t cnt = 0, ci = 0, g = 0;
    const uint32_t ring = desc_lo_sw128(smem_u32(smem));
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    for (int p = blockIdx.x; p < MT; p += gridDim.x) {
      for (int si = 0; si < a.n_steps; ++si, ++g) {
        const ChainK& s = a.s[si];
        const int KB = s.K / KBLK;
        cnt += KB;  // the ring stages that carried A
        mbar_wait(a_ready, g & 1);
        tc_fence_after();
        for (int n = 0; n < s.N; ++ci) {
          const int w = chunk_width(s, ci, n);
          const uint32_t ab = ci & 1, aph = (ci >> 1) & 1;
          mbar_wait(&acc_empty[ab], aph ^ 1);
          tc_fence_after();
          const uint32_t d = tmem + (ab ? s.acc1 : s.acc0);
          const uint32_t idesc = make_idesc(128, (w + 15) & ~15);
          for (int kb = 0; kb < KB; ++kb, ++cnt) {
            const uint32_t st_i = cnt % C_STAGES, ph = (cnt / C_STAGES) & 1;
            mbar_wait(&s_full[st_i], ph);
            tc_fence_after();
            const uint32_t bh0 = ring + st_i * 2 * BLK;
            const uint32_t ah = tmem + kb * 32, al = tmem + s.acol_lo + kb * 32;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k) {
                const uint32_t bh = bh0 + 2 * k;
                if (kb | k) umma_ts<true>(d, ah + 8 * k, bh, idesc); else umma_ts<false>(d, ah + 8 * k, bh, idesc);
                umma_ts<true>(d, al + 8 * k, bh, idesc);
                umma_ts<true>(d, ah + 8 * k, bh + BLK, idesc);
              }
              umma_commit(&s_empty[st_i]);
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&acc_full[ab]);
          __syncwarp();
          n += w;
        }
        if (elect_one()) umma_commit(a_free);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, hf = (warp - 2) >> 2;
    float* bias_s = reinterpret_cast<float*>(smem + C_OFF_BIAS) + (warp - 2) * 64;
    float* xp = reinterpret_cast<float*>(smem + C_OFF_XP) + (warp - 2) * 1024;
    uint32_t* xpu = reinterpret_cast<uint32_t*>(xp);
    const int xj = lane & 7;
    uint32_t cnt = 0, ci = 0, g = 0;
    for (int p = blockIdx.x; p < MT; p += gridDim.x) {
      const int m = p * TM + r;
      const bool row_ok = m < a.M;
      const long row0 = (long)p * TM + q * 32 + (lane >> 3);  // first of the 8 rows (stride 4) this lane stores
      for (int si = 0; si < a.n_steps; ++si, ++g) {
        const ChainK& s = a.s[si];
        const int KB = s.K / KBLK;
        {
          // ---- stage the panel of A into tensor memory: this thread copies row r of every other K block (128 bytes = 32 columns);
          // the two warps of a lane quarter (hf = 0 / 1) take the even / odd K blocks
          if (g > 0) mbar_wait(a_free, (g - 1) & 1);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb, ++cnt) {
            if ((kb & 1) != hf) continue;
            const uint32_t st_i = cnt % C_STAGES, ph = (cnt / C_STAGES) & 1;
            mbar_wait(&s_full[st_i], ph);
            const unsigned char* st = smem + st_i * stage_bytes;
#pragma unroll
            for (int img = 0; img < 2; ++img) {
              uint32_t pk[32];
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 x = *reinterpret_cast<const uint4*>(st + img * TILE_BYTES + sw128_offset(r, c8 * 8));
                pk[4 * c8] = x.x; pk[4 * c8 + 1] = x.y; pk[4 * c8 + 2] = x.z; pk[4 * c8 + 3] = x.w;
              }
              tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (img ? s.acol_lo : 0) + kb * 32, pk);
            }
            named_bar_sync(hf ? 3 : 1, 128);  // all four warps staging this K block have read it
            if (threadIdx.x == 64 + hf * 128) mbar_arrive(&s_empty[st_i]);
          }
          tc_fence_before();
          mbar_arrive(a_ready);
        }
        const float post = (row_ok && s.row_post) ? s.row_post[m] : 1.f;
        for (int n = 0; n < s.N; ++ci) {
          const int w = chunk_width(s, ci, n);
          const uint32_t ab = ci & 1, aph = (ci >> 1) & 1;
          const int nw0 = n + hf * 64;        // first column of this warp's 64-column share of the chunk
          const bool has_cols = hf * 64 < w;  // warp-uniform
          if (has_cols && s.bias) {
            float bv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) bv[u] = (nw0 + u * 32 + lane < s.N) ? __ldg(s.bias + nw0 + u * 32 + lane) : 0.f;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 2; ++u) bias_s[u * 32 + lane] = bv[u];
            __syncwarp();
          }
          const uint32_t taddr = tmem + (ab ? s.acc1 : s.acc0) + hf * 64 + ((uint32_t)(q * 32) << 16);
          // what leaves the step is dispatched once per chunk to a compile-time-specialised epilogue (see gemm_tc.cu for why)
          auto run = [&](auto epi_c) {
            constexpr int EPI = decltype(epi_c)::value;
            constexpr bool kC = (EPI & 1) != 0, kHi = (EPI & 2) != 0, kLo = (EPI & 4) != 0, kRes = (EPI & 8) != 0, kWide = (EPI & 16) != 0;
            float4 rv[8];
            auto fetch_res = [&](int nc) {  // residual of the 32-column sub-chunk starting at column nc, coalesced layout
              const int nn = nc + xj * 4;
              const float* rp = s.res + row0 * s.ldres + nn;
#pragma unroll
              for (int i8 = 0; i8 < 8; ++i8, rp += 4 * s.ldres)
                rv[i8] = (row0 + 4 * i8 < a.M && nn < s.N) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            if constexpr (kRes) { if (has_cols) fetch_res(nw0); }
            mbar_wait(&acc_full[ab], aph);
            tc_fence_after();
            if (!has_cols) {
              tc_fence_before();
              mbar_arrive(&acc_empty[ab]);
            } else if constexpr (kWide) {
              float v[64];
              tmem_ld32_issue(taddr, v);
              tmem_ld32_issue(taddr + 32, v + 32);
              tmem_wait_ld();
              tc_fence_before();
              mbar_arrive(&acc_empty[ab]);
              if (s.bias) {
#pragma unroll
                for (int e = 0; e < 64; e += 4) {
                  const float4 bv = *reinterpret_cast<const float4*>(bias_s + e);
                  v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
                }
              }
              if (s.relu) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] = fmaxf(v[e], 0.f);
              }
              if (s.row_post) {
#pragma unroll
                for (int e = 0; e < 64; ++e) v[e] *= post;
              }
              const long ooff = row0 * s.ldo + nw0 + xj * 8;
#pragma unroll
              for (int img = 0; img < (kLo ? 2 : 1); ++img) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  uint32_t pk[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float x0 = v[8 * j + 2 * u], x1 = v[8 * j + 2 * u + 1];
                    pk[u] = img == 0 ? pack_bf16(x0, x1) : pack_bf16(x0 - bf16_round(x0), x1 - bf16_round(x1));
                  }
                  *reinterpret_cast<uint4*>(xpu + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                __syncwarp();
                bf16* op = (img == 0 ? s.out_hi : s.out_lo) + ooff;
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8, op += 4 * s.ldo) {
                  const uint4 x = *reinterpret_cast<const uint4*>(xpu + (lane >> 3) * 32 + i8 * 128 + ((xj ^ (((i8 & 1) << 2) | (lane >> 3))) << 2));
                  if (row0 + 4 * i8 < a.M) *reinterpret_cast<uint4*>(op) = x;
                }
                __syncwarp();
              }
            } else {
#pragma unroll 1
              for (int c0 = 0; c0 < 64; c0 += 32) {
                const int n0 = nw0 + c0;
                if (hf * 64 + c0 >= w) break;  // warp-uniform
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (c0 == 32 || hf * 64 + 32 >= w) {  // last read of this accumulator by this warp
                  tc_fence_before();
                  mbar_arrive(&acc_empty[ab]);
                }
                if (s.bias) {
#pragma unroll
                  for (int e = 0; e < 32; e += 4) {
                    const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + e);
                    v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
                  }
                }
                if (s.relu) {
#pragma unroll
                  for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
                }
                if (s.row_post) {
#pragma unroll
                  for (int e = 0; e < 32; ++e) v[e] *= post;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(xp + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                const int nn = n0 + xj * 4;
                const bool col_ok = nn < s.N;
                const float* xr = xp + (lane >> 3) * 32;
                float* cp = kC ? s.C + row0 * s.ldc + nn : nullptr;
                bf16* hp = kHi ? s.out_hi + row0 * s.ldo + nn : nullptr;
                bf16* lp = kLo ? s.out_lo + row0 * s.ldo + nn : nullptr;
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                  float4 x = *reinterpret_cast<const float4*>(xr + i8 * 128 + ((xj ^ (((i8 & 1) << 2) | (lane >> 3))) << 2));
                  if constexpr (kRes) { x.x += rv[i8].x; x.y += rv[i8].y; x.z += rv[i8].z; x.w += rv[i8].w; }
                  if (col_ok && row0 + 4 * i8 < a.M) {
                    if constexpr (kC) *reinterpret_cast<float4*>(cp) = x;
                    if constexpr (kHi) *reinterpret_cast<uint2*>(hp) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
                    if constexpr (kLo)
                      *reinterpret_cast<uint2*>(lp) = make_uint2(pack_bf16(x.x - bf16_round(x.x), x.y - bf16_round(x.y)),
                                                                pack_bf16(x.z - bf16_round(x.z), x.w - bf16_round(x.w)));
                  }
                  if constexpr (kC) cp += 4 * s.ldc;
                  if constexpr (kHi) hp += 4 * s.ldo;
                  if constexpr (kLo) lp += 4 * s.ldo;
                }
                __syncwarp();
                if constexpr (kRes) { if (c0 == 0 && hf * 64 + 32 < w) fetch_res(n0 + 32); }
              }
            }
          };
          switch (s.epi) {  // warp-uniform
            case 1: run(std::integral_constant<int, 1>{}); break;
            case 7: run(std::integral_constant<int, 7>{}); break;
            case 9: run(std::integral_constant<int, 9>{}); break;
            case 15: run(std::integral_constant<int, 15>{}); break;
            case 18: run(std::integral_constant<int, 18>{}); break;
            default: run(std::integral_constant<int, 22>{}); break;
          }
          n += w;
          cnt += KB;  // the ring stages that carried this chunk's weight blocks (the staging warps index the ring by cnt)
        }
        if (s.ln_w) {
          // LayerNorm of the step's output rows: every chunk of the panel has been stored (by all eight warps) after the barrier
          named_bar_sync(2, 256);
          const int ew = warp - 2;
#pragma unroll 1
          for (int rr = 0; rr < 16; rr += 4) {  // 4 rows per call: their loads are in flight together (row_ops.cuh)
            const long mr = (long)p * TM + ew * 16 + rr;
            const int nv = (int)min((long)4, (long)a.M - mr);
            if (nv <= 0) break;
            const float* sc = s.ln_scale ? s.ln_scale + mr : nullptr;
            bf16* yh = s.ln_hi ? s.ln_hi + mr * s.ld_ln : nullptr;
            bf16* yl = s.ln_hi ? s.ln_lo + mr * s.ld_ln : nullptr;
            if (s.N == 256) layernorm_rows<256, 4>(s.C + mr * s.ldc, nullptr, s.ldc, s.ln_w, s.ln_b, sc, s.ln_out + mr * s.ld_ln, s.ld_ln, yh, yl, nullptr, s.ld_ln, lane, nv);
            else layernorm_rows<320, 4>(s.C + mr * s.ldc, nullptr, s.ldc, s.ln_w, s.ln_b, sc, s.ln_out + mr * s.ld_ln, s.ld_ln, yh, yl, nullptr, s.ld_ln, lane, nv);
          }
        }
        // this thread's global stores of the step become visible to the async proxy (the next step's TMA loads) before it arrives
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_arrive(step_done);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

void gemm_chain(const ChainStep* steps, int n_steps, int M, cudaStream_t st) {
  S2S_CHECK(n_steps >= 1 && n_steps <= CHAIN_MAX_STEPS && M > 0, "gemm_chain: bad step count");
  ChainMaps maps;
  ChainKArgs k;
  memset(&maps, 0, sizeof(maps));
  memset(&k, 0, sizeof(k));
  k.M = M; k.n_steps = n_steps;
  for (int i = 0; i < n_steps; ++i) {
    const ChainStep& t = steps[i];
    S2S_CHECK(t.K % 64 == 0 && t.K > 0 && t.K <= 320 && t.N > 0 && t.N % 4 == 0, "gemm_chain: K must be a multiple of 64 up to 320, N a multiple of 4");
    S2S_CHECK(t.A_hi && t.A_lo && t.W_hi && t.W_lo, "gemm_chain: split-bf16 operands missing");
    const long lda = t.lda ? t.lda : t.K;
    maps.m[i][0] = make_bf16_2d_map(t.A_hi, M, t.K, lda);
    maps.m[i][1] = make_bf16_2d_map(t.A_lo, M, t.K, lda);
    maps.m[i][2] = make_bf16_2d_map(t.W_hi, t.N, t.K, t.ldw);
    maps.m[i][3] = make_bf16_2d_map(t.W_lo, t.N, t.K, t.ldw);
    ChainK& s = k.s[i];
    s.N = t.N; s.K = t.K; s.relu = t.relu;
    s.acol_lo = t.K / 2; s.acc0 = t.K; s.acc1 = t.K + 128; s.w1 = 512 - t.K - 128 >= 128 ? 128 : 64;
    s.bias = t.bias; s.res = t.res; s.row_post = t.row_post; s.C = t.C; s.out_hi = t.out_hi; s.out_lo = t.out_lo;
    s.ldc = t.ldc; s.ldres = t.ldres; s.ldo = t.ldo;
    int epi = (t.C ? 1 : 0) | (t.out_hi ? 2 : 0) | (t.out_lo ? 4 : 0) | (t.res ? 8 : 0);
    S2S_CHECK(!(epi & 4) || (epi & 2), "gemm_chain: a lo image needs a hi image");
    S2S_CHECK((!t.C || t.ldc % 4 == 0) && (!t.res || t.ldres % 4 == 0) && (!t.out_hi || t.ldo % 4 == 0), "gemm_chain: pitches must be multiples of 4");
    if ((epi == 2 || epi == 6) && t.N % 64 == 0 && t.ldo % 8 == 0) epi |= 16;
    S2S_CHECK(epi == 1 || epi == 7 || epi == 9 || epi == 15 || epi == 18 || epi == 22, "gemm_chain: unsupported output combination " + std::to_string(epi));
    s.epi = epi;
    if (t.ln_w) {
      S2S_CHECK(t.C && t.ln_b && t.ln_out && (t.N == 256 || t.N == 320) && t.ld_ln % 4 == 0 && t.ld_ln >= t.N && (!t.ln_hi || t.ln_lo),
                "gemm_chain: LayerNorm needs an fp32 output of 256 or 320 columns");
      s.ln_w = t.ln_w; s.ln_b = t.ln_b; s.ln_scale = t.ln_scale; s.ln_out = t.ln_out; s.ln_hi = t.ln_hi; s.ln_lo = t.ln_lo; s.ld_ln = t.ld_ln;
    }
  }
  static bool configured = false;
  if (!configured) {
    S2S_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C_SMEM));
    configured = true;
  }
  S2S_PROF(g_profile_on ? prof_intern("gemm_chain x" + std::to_string(n_steps)) : "gemm_chain", st);
  const int panels = ceil_div(M, TM);
  launch_pdl(gemm_chain_kernel, panels < sm_count() ? panels : sm_count(), C_THREADS, C_SMEM, st, maps, k);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
