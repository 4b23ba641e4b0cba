// Tensor-core GEMM for the node track:  C = epilogue(A * B^T)  with A [M][K] and B [N][K] bf16, K-major, fetched by
// TMA tensor maps into SWIZZLE_128B operand blocks, tcgen05.mma (128x128x16) into double-buffered TMEM accumulators,
// epilogue from TMEM by eight warps (the epilogue is issue-bound per warp: ncu shows one warp per scheduler stalling on
// its own dependent instructions, so two warps per scheduler halve its time).  Persistent CTAs loop over 128x128 output tiles.
//
// passes = 1: plain bf16 product (IPA q/k/v projections, q.k^T, P.V — tools/precision_probe.py shows these tolerate it)
// passes = 3: split-bf16 product  A_hi B_hi + A_lo B_hi + A_hi B_lo  accumulated in fp32 (≈16 mantissa bits per operand)
//             for the layers of the residual stream that do not tolerate single bf16 rounding.
#include <algorithm>
#include <cstdlib>

#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

constexpr int G_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, 64 columns each)
constexpr int G_SMEM_RING = 12 * TILE_BYTES;  // 192 KiB of operand blocks: 6 stages x 2 blocks, or 3 stages x 4 blocks
constexpr int G_OFF_BAR = G_SMEM_RING;
constexpr int G_OFF_BIAS = G_OFF_BAR + 32 * 8 + 16;  // per-epilogue-warp bias slice of the current tile, [8][64] fp32
constexpr int G_OFF_XP = G_OFF_BIAS + 8 * 64 * 4;    // per-epilogue-warp 32 x 32 fp32 transposition buffer (see the epilogue)
constexpr int G_SMEM = G_OFF_XP + 8 * 32 * 32 * 4;   // 231 696 B of the 232 448 B a CTA can have

struct TcKernelArgs {
  int a_cb, a_ch, a_rb, a_rh;  // A box coordinates: col = ib*a_cb + ih*a_ch + kb*64, row = ib*a_rb + ih*a_rh + m0
  int b_cb, b_ch, b_rb, b_rh;
  int M, N, K, nb, nh, passes, relu, vt_L;
  int vt_col0, vt_stride, vt_off, vt_width, vt_heads;  // which output columns are 'v' and how they group into heads
  float alpha;
  const float *bias, *row_pre, *row_post, *res;
  float* C;
  long ldc, sCb, sCh, ldres;
  bf16 *out_hi, *out_lo;  // optional dense bf16 copies of the result, row pitch ldo (non-batched calls only)
  long ldo;
  bf16 *out_vt, *out_vt_lo;  // optional transposed bf16 (hi, lo) copy of the v columns of a fused q|k|v projection
  int dbg;                // timing experiments only (S2S_GEMM_DEBUG): 8 no TMA, 16 no MMA, 32 no epilogue
  // second K segment (passes == 1 only): K2 more reduction columns fetched through the mAl / mBl maps
  int K2, a2_cb, a2_ch, a2_rb, a2_rh, b2_cb, b2_ch, b2_rb, b2_rh;
  long bias_sb, bias_sh;  // batch strides of `bias` (0: one bias vector for every batch)
  int coalesced;          // 1: outputs / residual go through the per-warp transposition buffer (all pitches and N % 4 == 0)
  int b_mn;               // B operand is MN-major ([K][N] source), see TcGemm
  int n_split, n_per;     // panel kernel: CTAs per row panel and output columns per CTA (a multiple of 64), see gemm_tc()
};

// EPI >= 0: coalesced epilogue specialised at compile time on what leaves the kernel (bit 0: fp32 C, 1: bf16 hi image,
// 2: bf16 lo image, 3: residual added) — the generic epilogue spent ~800 instructions per 32 x 32 chunk on option
// predicates and 64-bit address arithmetic (ncu: issue-bound at 44 % of the slots with 2.5 warps per scheduler).
// EPI < 0: generic row-layout epilogue (ragged N, unaligned pitches, transposed v output).
template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
               const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, TcKernelArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];  // SWIZZLE_128B operand blocks need 1024-byte alignment
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);
  uint64_t* s_full = bars;        // [6]
  uint64_t* s_empty = bars + 6;   // [6]
  uint64_t* acc_full = bars + 12;   // [2]
  uint64_t* acc_empty = bars + 14;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int blocks_per_stage = a.passes == 3 ? 4 : 2;
  const int n_stages = 12 / blocks_per_stage;
  const uint32_t stage_bytes = blocks_per_stage * TILE_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 6; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 32 * 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // barriers and tensor memory are set up: from here on the kernel reads what the previous one wrote
  const int MT = (a.M + TM - 1) / TM, NT = (a.N + 127) / 128, KB = (a.K + KBLK - 1) / KBLK;
  const int KB2 = (a.K2 + KBLK - 1) / KBLK, KBT = KB + KB2;
  const int n_tiles = a.nb * a.nh * MT * NT;
  constexpr uint32_t IDESC = make_idesc(128, 128);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % NT, mt = (tile / NT) % MT, bz = tile / (NT * MT);
        const int ib = bz / a.nh, ih = bz % a.nh;
        const int arow = ib * a.a_rb + ih * a.a_rh + mt * TM, acol = ib * a.a_cb + ih * a.a_ch;
        const int brow = ib * a.b_rb + ih * a.b_rh + nt * 128, bcol = ib * a.b_cb + ih * a.b_ch;
        for (int kb = 0; kb < KBT; ++kb, ++cnt) {
          const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
          mbar_wait(&s_empty[s], ph ^ 1);
          if (a.dbg & 8) { mbar_arrive(&s_full[s]); continue; }
          mbar_expect_tx(&s_full[s], stage_bytes);
          unsigned char* st = smem + s * stage_bytes;
          if (kb >= KB) {  // second K segment
            const int k2 = (kb - KB) * KBLK;
            tma_load_2d(st, &mAl, ib * a.a2_cb + ih * a.a2_ch + k2, ib * a.a2_rb + ih * a.a2_rh + mt * TM, &s_full[s]);
            tma_load_2d(st + TILE_BYTES, &mBl, ib * a.b2_cb + ih * a.b2_ch + k2, ib * a.b2_rb + ih * a.b2_rh + nt * 128, &s_full[s]);
            continue;
          }
          tma_load_2d(st, &mAh, acol + kb * KBLK, arow, &s_full[s]);
          if (a.b_mn) {  // two [64 k-rows x 64 n] boxes = the 128 n columns of this tile for one 64-wide K block
            const int r0 = ib * a.b_rb + ih * a.b_rh + kb * KBLK, c0 = bcol + nt * 128;
            tma_load_2d(st + TILE_BYTES, &mBh, c0, r0, &s_full[s]);
            tma_load_2d(st + TILE_BYTES + TILE_BYTES / 2, &mBh, c0 + 64, r0, &s_full[s]);
            if (a.passes == 3) {
              tma_load_2d(st + 2 * TILE_BYTES, &mAl, acol + kb * KBLK, arow, &s_full[s]);
              tma_load_2d(st + 3 * TILE_BYTES, &mBl, c0, r0, &s_full[s]);
              tma_load_2d(st + 3 * TILE_BYTES + TILE_BYTES / 2, &mBl, c0 + 64, r0, &s_full[s]);
            }
            continue;
          }
          tma_load_2d(st + TILE_BYTES, &mBh, bcol + kb * KBLK, brow, &s_full[s]);
          if (a.passes == 3) {
            tma_load_2d(st + 2 * TILE_BYTES, &mAl, acol + kb * KBLK, arow, &s_full[s]);
            tma_load_2d(st + 3 * TILE_BYTES, &mBl, bcol + kb * KBLK, brow, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp runs the loop with warp-uniform values; one elected lane issues (see tc_common.cuh)
    uint32_t cnt = 0, it = 0;
    const uint32_t ring = desc_lo_sw128(smem_u32(smem));
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    // dbg bits 64 / 128 (timing experiments only, results are garbage): issue every MMA as N = 256 / N = 64
    const uint32_t idesc_n = (a.dbg & 64) ? make_idesc(128, 256) : (a.dbg & 128) ? make_idesc(128, 64) : IDESC;
    const uint32_t idesc = a.b_mn ? (idesc_n | (1u << 16)) : idesc_n;        // bit 16: B is MN-major
    const uint32_t bmn_fix = ((uint32_t)(TILE_BYTES / 2) >> 4 << 16) - (1u << 16);  // LBO field 1 -> 512 (8 KB)
    const uint32_t stage_units = blocks_per_stage * BLK;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1, aph = (it >> 1) & 1;
      mbar_wait(&acc_empty[ab], aph ^ 1);
      tc_fence_after();
      const uint32_t d = (a.dbg & 64) ? tmem : tmem + ab * 128;
      for (int kb = 0; kb < KBT; ++kb, ++cnt) {
        const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
        mbar_wait(&s_full[s], ph);
        tc_fence_after();
        const uint32_t st = ring + s * stage_units;
        const int ksteps = (min(KBLK, kb < KB ? a.K - kb * KBLK : a.K2 - (kb - KB) * KBLK) + 15) / 16;
        if (elect_one()) {
          if (!(a.dbg & 16)) {
            for (int k = 0; k < ksteps; ++k) {
              // K-major B: k-step = +32 B inside the 128-byte rows.  MN-major B: k-step = 16 rows = +2 KB; the descriptor's
              // leading-byte-offset field (bits 16..29) carries the 8 KB distance between the two 64-column halves.
              const uint32_t ah = st + 2 * k;
              const uint32_t bh = a.b_mn ? st + BLK + 128 * k + bmn_fix : st + BLK + 2 * k;
              if (a.dbg & 256) {  // timing experiment only: consecutive MMAs alternate between the two accumulator buffers
                umma_ss<true>(tmem + (k & 1) * 128, ah, bh, idesc);
                continue;
              }
              if (kb | k) umma_ss<true>(d, ah, bh, idesc); else umma_ss<false>(d, ah, bh, idesc);
              if (a.passes == 3) {
                umma_ss<true>(d, st + 2 * BLK + 2 * k, bh, idesc);
                umma_ss<true>(d, ah, bh + 2 * BLK, idesc);
              }
            }
          }
          umma_commit(&s_empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[ab]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane, hf = (warp - 2) >> 2;
    float* bias_s = reinterpret_cast<float*>(smem + G_OFF_BIAS) + (warp - 2) * 64;
    // Coalescing: tcgen05.ld hands every thread one accumulator ROW, so storing from that layout makes each 16-byte store
    // of a warp touch 32 different 128-byte lines (ncu: 32 L1 tag requests / wavefronts per STG.128, and the LSU data pipe,
    // not the tensor pipe, bounded the node-track GEMMs).  The finished 32 x 32 chunk is therefore turned through shared
    // memory (16-byte chunks XOR-swizzled by row: conflict-free both ways) and leaves as rows of 128 contiguous bytes.
    float* xp = reinterpret_cast<float*>(smem + G_OFF_XP) + (warp - 2) * 1024;
    constexpr int cw = 64;  // columns of a tile per epilogue warp
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % NT, mt = (tile / NT) % MT, bz = tile / (NT * MT);
      const int ib = bz / a.nh, ih = bz % a.nh;
      const uint32_t ab = it & 1, aph = (it >> 1) & 1;
      const int m = mt * TM + r;
      const bool row_ok = m < a.M;
      const float pre = (row_ok && a.row_pre) ? a.row_pre[m] * a.alpha : a.alpha;
      const float post = (row_ok && a.row_post) ? a.row_post[m] : 1.f;
      float* crow = a.C ? a.C + ib * a.sCb + ih * a.sCh + (long)m * a.ldc : nullptr;
      const float* rrow = a.res ? a.res + ib * a.sCb + ih * a.sCh + (long)m * a.ldres : nullptr;
      const float* bias = a.bias ? a.bias + ib * a.bias_sb + ih * a.bias_sh : nullptr;
      if (bias) {
        // This warp's 64 bias values go to shared memory BEFORE the accumulator wait: a global load issued inside the
        // chunk loop queues behind the previous chunk's scattered stores and its latency is fully exposed (ncu: the
        // dependent FADD was the top long-scoreboard stall of the kernel).
        const int nb0 = nt * 128 + hf * cw + lane;
        float bv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) bv[u] = (nb0 + u * 32 < a.N) ? __ldg(bias + nb0 + u * 32) : 0.f;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 2; ++u) bias_s[u * 32 + lane] = bv[u];
        __syncwarp();
      }
      // Residual values of a chunk are fetched (in the coalesced layout: lane -> row it*4 + lane/8, columns 4*(lane%8)..+3)
      // before the accumulator wait / while the previous chunk drains, ahead of any store of this tile: `res` may alias `C`,
      // so loads placed after a store could not be hoisted by the compiler and each would expose a full L2 round trip.
      constexpr bool kC = EPI >= 0 && (EPI & 1), kHi = EPI >= 0 && (EPI & 2), kLo = EPI >= 0 && (EPI & 4), kRes = EPI >= 0 && (EPI & 8);
      const long boff = ib * a.sCb + ih * a.sCh;
      const int xj = lane & 7;
      const long row0 = (long)mt * TM + q * 32 + (lane >> 3);  // first of the 8 rows (stride 4) this lane stores
      float4 rv[8];
      auto fetch_res = [&](int c0) {
        const int n = nt * 128 + c0 + xj * 4;
        const float* rp = a.res + boff + row0 * a.ldres + n;
#pragma unroll
        for (int i8 = 0; i8 < 8; ++i8, rp += 4 * a.ldres)
          rv[i8] = (row0 + 4 * i8 < a.M && n < a.N) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if constexpr (kRes) fetch_res(hf * cw);
      mbar_wait(&acc_full[ab], aph);
      tc_fence_after();
      const uint32_t taddr = tmem + ab * 128 + ((uint32_t)(q * 32) << 16);
      if constexpr (EPI >= 16) {
        // bf16-only outputs (hi image, optionally lo; N % 64 == 0): the warp's 64 columns are packed to bf16 in the row layout
        // and turned through shared memory as whole 128-byte rows, so every 16-byte store of a warp covers four complete lines
        // (the 32-column fp32 route moves twice the bytes through shared memory and stores half lines).
        const int n0 = nt * 128 + hf * cw;
        if (n0 < a.N && !(a.dbg & 32)) {
          float v[64];
          tmem_ld32_issue(taddr + hf * cw, v);
          tmem_ld32_issue(taddr + hf * cw + 32, v + 32);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);  // the accumulator is in registers: the MMA warp may refill it
          if (pre != 1.f) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] *= pre;
          }
          if (bias) {
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + e);
              v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
            }
          }
          if (a.relu) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (a.row_post) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] *= post;
          }
          uint32_t* xpu = reinterpret_cast<uint32_t*>(xp);
          const long ooff = boff + row0 * a.ldo + n0 + xj * 8;
#pragma unroll
          for (int img = 0; img < ((EPI & 4) ? 2 : 1); ++img) {
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
            bf16* op = (img == 0 ? a.out_hi : a.out_lo) + ooff;
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8, op += 4 * a.ldo) {
              const uint4 x = *reinterpret_cast<const uint4*>(xpu + (lane >> 3) * 32 + i8 * 128 + ((xj ^ (((i8 & 1) << 2) | (lane >> 3))) << 2));
              if (row0 + 4 * i8 < a.M) *reinterpret_cast<uint4*>(op) = x;
            }
            __syncwarp();
          }
        } else {
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = hf * cw; c0 < hf * cw + cw && !(a.dbg & 32); c0 += 32) {
        const int n0 = nt * 128 + c0;
        if (n0 >= a.N) break;  // warp-uniform
        float v[32];  // static indexing only below: stays in registers
        tmem_ld32(taddr + c0, v);
        if (EPI < 0 && !row_ok) continue;
        const bool full = n0 + 32 <= a.N;
        const bool vec_ok = full && (a.ldres & 3) == 0;
        if (pre != 1.f) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] *= pre;
        }
        if (bias) {
          const float* bs = bias_s + (c0 - hf * cw);
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bs + e);
            v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
          }
        }
        if (a.relu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
        }
        if (a.row_post) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] *= post;
        }
        if constexpr (EPI >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(xp + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const int n = n0 + xj * 4;
          const bool col_ok = n < a.N;  // N % 4 == 0: a group of 4 columns is inside or outside as a whole
          const float* xr = xp + (lane >> 3) * 32;
          float* cp = kC ? a.C + boff + row0 * a.ldc + n : nullptr;
          bf16* hp = kHi ? a.out_hi + boff + row0 * a.ldo + n : nullptr;
          bf16* lp = kLo ? a.out_lo + boff + row0 * a.ldo + n : nullptr;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            // row = i8*4 + lane/8, so row & 7 = ((i8 & 1) << 2) | (lane >> 3): the swizzle term is known per lane
            float4 x = *reinterpret_cast<const float4*>(xr + i8 * 128 + ((xj ^ (((i8 & 1) << 2) | (lane >> 3))) << 2));
            if constexpr (kRes) { x.x += rv[i8].x; x.y += rv[i8].y; x.z += rv[i8].z; x.w += rv[i8].w; }
            if (col_ok && row0 + 4 * i8 < a.M) {
              if constexpr (kC) *reinterpret_cast<float4*>(cp) = x;
              if constexpr (kHi) *reinterpret_cast<uint2*>(hp) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
              if constexpr (kLo)
                *reinterpret_cast<uint2*>(lp) = make_uint2(pack_bf16(x.x - bf16_round(x.x), x.y - bf16_round(x.y)),
                                                          pack_bf16(x.z - bf16_round(x.z), x.w - bf16_round(x.w)));
            }
            if constexpr (kC) cp += 4 * a.ldc;
            if constexpr (kHi) hp += 4 * a.ldo;
            if constexpr (kLo) lp += 4 * a.ldo;
          }
          __syncwarp();
          if constexpr (kRes) { if (c0 + 32 < hf * cw + cw) fetch_res(c0 + 32); }
          continue;
        }
        if (rrow) {
          if (vec_ok) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 rv = *reinterpret_cast<const float4*>(rrow + n0 + e);
              v[e] += rv.x; v[e + 1] += rv.y; v[e + 2] += rv.z; v[e + 3] += rv.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (n0 + e < a.N) v[e] += rrow[n0 + e];
          }
        }
        if (crow) {
          if (full && (a.ldc & 3) == 0) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(crow + n0 + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          } else {  // ragged last chunk (N = 80, 36, ...): whole groups of 4 columns still go out as one 16-byte store
            const bool vec = (a.ldc & 3) == 0;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              if (vec && n0 + e + 4 <= a.N) {
                *reinterpret_cast<float4*>(crow + n0 + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (n0 + e + u < a.N) crow[n0 + e + u] = v[e + u];
              }
            }
          }
        }
        if (a.out_hi) {
          const long ooff = ib * a.sCb + ih * a.sCh + (long)m * a.ldo + n0;
          bf16* hrow = a.out_hi + ooff;
          bf16* lrow = a.out_lo ? a.out_lo + ooff : nullptr;
          if (full && (a.ldo & 7) == 0) {
#pragma unroll
            for (int e = 0; e < 32; e += 8) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float x0 = v[e + 2 * u], x1 = v[e + 2 * u + 1];
                h[u] = pack_bf16(x0, x1);
                l[u] = pack_bf16(x0 - bf16_round(x0), x1 - bf16_round(x1));
              }
              *reinterpret_cast<uint4*>(hrow + e) = make_uint4(h[0], h[1], h[2], h[3]);
              if (lrow) *reinterpret_cast<uint4*>(lrow + e) = make_uint4(l[0], l[1], l[2], l[3]);
            }
          } else {
            const bool vec = (a.ldo & 3) == 0 && (a.sCh & 3) == 0 && (a.sCb & 3) == 0;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              if (vec && n0 + e + 4 <= a.N) {
                *reinterpret_cast<uint2*>(hrow + e) = make_uint2(pack_bf16(v[e], v[e + 1]), pack_bf16(v[e + 2], v[e + 3]));
                if (lrow)
                  *reinterpret_cast<uint2*>(lrow + e) = make_uint2(pack_bf16(v[e] - bf16_round(v[e]), v[e + 1] - bf16_round(v[e + 1])),
                                                                  pack_bf16(v[e + 2] - bf16_round(v[e + 2]), v[e + 3] - bf16_round(v[e + 3])));
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (n0 + e + u < a.N) {
                    const bf16 h = __float2bfloat16_rn(v[e + u]);
                    hrow[e + u] = h;
                    if (lrow) lrow[e + u] = __float2bfloat16_rn(v[e + u] - __bfloat162float(h));
                  }
                }
              }
            }
          }
        }
        if (a.out_vt && n0 + 32 > a.vt_col0) {
          // v columns, transposed per head: VT[((b*heads + h)*width + c)*L + j]  (row m = b*L + j; lanes = consecutive j)
          const int b = m / a.vt_L, j = m % a.vt_L;
          const int nn0 = n0 - a.vt_col0;  // >= 0: vt_col0 is a multiple of the 32-column chunk
          const int h0 = nn0 / a.vt_stride, w0 = nn0 % a.vt_stride - a.vt_off;
          const int hL = (nn0 + 31) / a.vt_stride, wL = (nn0 + 31) % a.vt_stride - a.vt_off;
          if (h0 == hL && (wL < 0 || w0 >= a.vt_width)) {
            // chunk lies entirely in the q/k part of one head: nothing to transpose (warp-uniform)
          } else if (full && h0 == hL && w0 >= 0 && wL < a.vt_width) {  // whole chunk inside one head's v range
            bf16* dh = a.out_vt + (((long)b * a.vt_heads + h0) * a.vt_width + w0) * a.vt_L + j;
            bf16* dl = a.out_vt_lo ? a.out_vt_lo + (dh - a.out_vt) : nullptr;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const bf16 hi = __float2bfloat16_rn(v[e]);
              dh[(long)e * a.vt_L] = hi;
              if (dl) dl[(long)e * a.vt_L] = __float2bfloat16_rn(v[e] - __bfloat162float(hi));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int nn = nn0 + e;
              if (nn < 0 || n0 + e >= a.N) continue;
              const int h = nn / a.vt_stride, c = nn % a.vt_stride - a.vt_off;
              if (c < 0 || c >= a.vt_width) continue;
              const long idx = (((long)b * a.vt_heads + h) * a.vt_width + c) * a.vt_L + j;
              const bf16 hi = __float2bfloat16_rn(v[e]);
              a.out_vt[idx] = hi;
              if (a.out_vt_lo) a.out_vt_lo[idx] = __float2bfloat16_rn(v[e] - __bfloat162float(hi));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// =====================================================================================================================
// Panel GEMM: A resident in tensor memory (TS-mode MMAs), weights streamed.
//
// Why: with both operands in shared memory (SS mode) every tcgen05.mma of the kernel above costs ~210 cycles + 0.42 N
// (measured, S2S_GEMM_DEBUG 40 / 104 / 168: MMA-only time 87 / 97 / 117 us for N = 64 / 128 / 256 at identical instruction
// counts), i.e. a 128x128x16 MMA takes ~265 cycles against a 64-cycle tensor floor, and the node-track GEMMs — all
// K <= 320 with 3 split-bf16 passes — were bound by that per-instruction cost, not by FLOPs, bytes or the epilogue.
// The fused EdgeTransition kernel, whose A operand lives in tensor memory, issues the same MMAs at ~80 cycles.
//
// So for the non-batched GEMMs (C = act(A W^T + b), A [M, K] = activations): one CTA owns a 128-row panel of A, copies it
// ONCE into tensor memory (TMA -> SW128 shared block -> registers -> tcgen05.st; hi and lo images) and then walks all N
// output columns in chunks of <= 128, streaming only the weight blocks through the shared-memory ring.  A_hi serves two
// of the three split-bf16 passes; A is read from L2 once per panel instead of once per 128 output columns.
// Tensor memory: [A_hi | A_lo] (K/2 columns each) + two accumulator buffers (128 + 128, or 128 + 64 when K = 320 leaves
// only 192 columns).  Roles: TMA producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter) that also stage A at
// the start of a panel, the two warps of a quarter taking the even / odd K blocks.  Coalesced compile-time-specialised
// epilogue as above (EPI >= 0 only).  With few panels the output columns of a panel are split over several CTAs (n_split).
struct PanelGeom {
  int acol_lo;     // first TMEM column of A_lo (A_hi starts at column 0)
  int acc0, acc1;  // first TMEM column of the two accumulator buffers
  int w1;          // width of buffer 1 (buffer 0 is 128 wide)
};

template <int EPI>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_panel_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                  const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, TcKernelArgs a, PanelGeom pg) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);
  uint64_t* s_full = bars;          // [12]
  uint64_t* s_empty = bars + 12;    // [12]
  uint64_t* acc_full = bars + 24;   // [2]
  uint64_t* acc_empty = bars + 26;  // [2]
  uint64_t* a_ready = bars + 28;    // A of the current panel is in tensor memory
  uint64_t* a_free = bars + 29;     // every MMA of the panel has completed: A may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int bps = a.passes == 3 ? 2 : 1;  // blocks per ring stage: (hi, lo) or hi
  const int n_stages = 12 / bps;
  const uint32_t stage_bytes = bps * TILE_BYTES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 12; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 32 * 8);
    }
    mbar_init(a_ready, 32 * 8);
    mbar_init(a_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // barriers and tensor memory are set up: from here on the kernel reads what the previous one wrote
  const int MT = (a.M + TM - 1) / TM, KB = a.K / KBLK;
  const int KB2 = (a.K2 + KBLK - 1) / KBLK, KBT = KB + KB2;  // second K segment (passes == 1): through the mAl / mBl maps
  const int NP = a.nb * a.nh * MT;                            // panels: (batch, head, 128-row tile)
  auto ksteps_of = [&](int kb) { return (min(KBLK, kb < KB ? a.K - kb * KBLK : a.K2 - (kb - KB) * KBLK) + 15) / 16; };
  // chunk ci of the whole CTA run uses accumulator buffer ci & 1; its width follows from the buffer and what is left of N
  auto chunk_width = [&](uint32_t ci, int n_done, int n_hi) { return min((ci & 1) ? pg.w1 : 128, n_hi - n_done); };
  // With fewer panels than SMs the output columns of a panel are split over n_split CTAs (work item pp = panel * n_split + part):
  // each stages the panel once and walks its own column range [n_lo, n_hi), so a small-M GEMM is not bound by one SM streaming
  // the whole weight matrix and issuing every MMA.
  const int NW = NP * a.n_split;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0, ci = 0;
      for (int pp = blockIdx.x; pp < NW; pp += gridDim.x) {
        const int p = pp / a.n_split, n_lo = (pp % a.n_split) * a.n_per, n_hi = min(a.N, n_lo + a.n_per);
        const int mt = p % MT, bz = p / MT, ib = bz / a.nh, ih = bz % a.nh;
        const int arow = ib * a.a_rb + ih * a.a_rh + mt * TM, acol = ib * a.a_cb + ih * a.a_ch;
        const int brow = ib * a.b_rb + ih * a.b_rh, bcol = ib * a.b_cb + ih * a.b_ch;
        for (int kb = 0; kb < KBT; ++kb, ++cnt) {  // the panel of A (consumed by the staging warps)
          const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
          mbar_wait(&s_empty[s], ph ^ 1);
          mbar_expect_tx(&s_full[s], stage_bytes);
          unsigned char* st = smem + s * stage_bytes;
          if (kb >= KB) {
            tma_load_2d(st, &mAl, ib * a.a2_cb + ih * a.a2_ch + (kb - KB) * KBLK, ib * a.a2_rb + ih * a.a2_rh + mt * TM, &s_full[s]);
            continue;
          }
          tma_load_2d(st, &mAh, acol + kb * KBLK, arow, &s_full[s]);
          if (bps == 2) tma_load_2d(st + TILE_BYTES, &mAl, acol + kb * KBLK, arow, &s_full[s]);
        }
        for (int n = n_lo; n < n_hi; ++ci) {  // weight blocks: 128 rows of W from row n (rows past the end read as zeros), one K block
          const int w = chunk_width(ci, n, n_hi);
          for (int kb = 0; kb < KBT; ++kb, ++cnt) {
            const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
            mbar_wait(&s_empty[s], ph ^ 1);
            mbar_expect_tx(&s_full[s], stage_bytes);
            unsigned char* st = smem + s * stage_bytes;
            if (kb >= KB) {
              tma_load_2d(st, &mBl, ib * a.b2_cb + ih * a.b2_ch + (kb - KB) * KBLK, ib * a.b2_rb + ih * a.b2_rh + n, &s_full[s]);
            } else if (a.b_mn) {  // B given as [K rows][N columns]: two [64 k x 64 n] boxes per K block
              const int r0 = brow + kb * KBLK, c0 = bcol + n;
              tma_load_2d(st, &mBh, c0, r0, &s_full[s]);
              tma_load_2d(st + TILE_BYTES / 2, &mBh, c0 + 64, r0, &s_full[s]);
              if (bps == 2) {
                tma_load_2d(st + TILE_BYTES, &mBl, c0, r0, &s_full[s]);
                tma_load_2d(st + TILE_BYTES + TILE_BYTES / 2, &mBl, c0 + 64, r0, &s_full[s]);
              }
            } else {
              tma_load_2d(st, &mBh, bcol + kb * KBLK, brow + n, &s_full[s]);
              if (bps == 2) tma_load_2d(st + TILE_BYTES, &mBl, bcol + kb * KBLK, brow + n, &s_full[s]);
            }
          }
          n += w;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (whole warp converged, one elected lane issues): D[acc] += A[tmem] * W_blk^T
    uint32_t cnt = 0, ci = 0, pi = 0;
    const uint32_t ring = desc_lo_sw128(smem_u32(smem));
    constexpr uint32_t BLK = TILE_BYTES >> 4;
    const uint32_t stage_units = bps * BLK;
    const uint32_t bmn_fix = ((uint32_t)(TILE_BYTES / 2) >> 4 << 16) - (1u << 16);  // MN-major B: LBO field 1 -> 512 (8 KB between halves)
    for (int pp = blockIdx.x; pp < NW; pp += gridDim.x, ++pi) {
      const int n_lo = (pp % a.n_split) * a.n_per, n_hi = min(a.N, n_lo + a.n_per);
      cnt += KBT;  // the ring stages that carried A
      mbar_wait(a_ready, pi & 1);
      tc_fence_after();
      for (int n = n_lo; n < n_hi; ++ci) {
        const int w = chunk_width(ci, n, n_hi);
        const uint32_t ab = ci & 1, aph = (ci >> 1) & 1;
        mbar_wait(&acc_empty[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + (ab ? pg.acc1 : pg.acc0);
        const uint32_t idesc = make_idesc(128, (w + 15) & ~15) | (a.b_mn ? (1u << 16) : 0u);  // bit 16: B is MN-major
        for (int kb = 0; kb < KBT; ++kb, ++cnt) {
          const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
          mbar_wait(&s_full[s], ph);
          tc_fence_after();
          const uint32_t bh0 = ring + s * stage_units;
          const uint32_t ah = tmem + kb * 32, al = tmem + pg.acol_lo + kb * 32;
          const int ksteps = ksteps_of(kb);
          if (elect_one()) {
            if (!(a.dbg & 16)) {
              for (int k = 0; k < ksteps; ++k) {
                // K-major B: k-step = +32 B inside the 128-byte rows; MN-major B: k-step = 16 rows = +2 KB
                const uint32_t bh = a.b_mn ? bh0 + 128 * k + bmn_fix : bh0 + 2 * k;
                if (kb | k) umma_ts<true>(d, ah + 8 * k, bh, idesc); else umma_ts<false>(d, ah + 8 * k, bh, idesc);
                if (bps == 2) {
                  umma_ts<true>(d, al + 8 * k, bh, idesc);
                  umma_ts<true>(d, ah + 8 * k, bh + BLK, idesc);
                }
              }
            }
            umma_commit(&s_empty[s]);
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
  } else {
    const int q = warp & 3, r = q * 32 + lane, hf = (warp - 2) >> 2;
    float* bias_s = reinterpret_cast<float*>(smem + G_OFF_BIAS) + (warp - 2) * 64;
    float* xp = reinterpret_cast<float*>(smem + G_OFF_XP) + (warp - 2) * 1024;
    uint32_t* xpu = reinterpret_cast<uint32_t*>(xp);
    constexpr bool kC = (EPI & 1) != 0, kHi = (EPI & 2) != 0, kLo = (EPI & 4) != 0, kRes = (EPI & 8) != 0, kWide = (EPI & 16) != 0;
    const int xj = lane & 7;
    uint32_t cnt = 0, ci = 0, pi = 0;
    for (int pp = blockIdx.x; pp < NW; pp += gridDim.x, ++pi) {
      const int p = pp / a.n_split, n_lo = (pp % a.n_split) * a.n_per, n_hi = min(a.N, n_lo + a.n_per);
      const int mt = p % MT, bz = p / MT, ib = bz / a.nh, ih = bz % a.nh;
      const long boff = ib * a.sCb + ih * a.sCh;
      const float* bias = a.bias ? a.bias + ib * a.bias_sb + ih * a.bias_sh : nullptr;
      {
        // ---- stage the panel of A into tensor memory: this thread copies row r of every other K block (128 bytes = 32 columns);
        // the two warps of a lane quarter (hf = 0 / 1) take the even / odd K blocks
        if (pi > 0) mbar_wait(a_free, (pi - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < KBT; ++kb, ++cnt) {
          if ((kb & 1) != hf) continue;
          const uint32_t s = cnt % n_stages, ph = (cnt / n_stages) & 1;
          mbar_wait(&s_full[s], ph);
          const unsigned char* st = smem + s * stage_bytes;
          for (int img = 0; img < bps; ++img) {
            uint32_t pk[32];
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 x = *reinterpret_cast<const uint4*>(st + img * TILE_BYTES + sw128_offset(r, c8 * 8));
              pk[4 * c8] = x.x; pk[4 * c8 + 1] = x.y; pk[4 * c8 + 2] = x.z; pk[4 * c8 + 3] = x.w;
            }
            tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + (img ? pg.acol_lo : 0) + kb * 32, pk);
          }
          named_bar_sync(1 + hf, 128);  // all four warps staging this K block have read it
          if (threadIdx.x == 64 + hf * 128) mbar_arrive(&s_empty[s]);
        }
        tc_fence_before();
        mbar_arrive(a_ready);
      }
      const int m = mt * TM + r;
      const bool row_ok = m < a.M;
      const float pre = (row_ok && a.row_pre) ? a.row_pre[m] * a.alpha : a.alpha;
      const float post = (row_ok && a.row_post) ? a.row_post[m] : 1.f;
      const long row0 = (long)mt * TM + q * 32 + (lane >> 3);  // first of the 8 rows (stride 4) this lane stores
      for (int n = n_lo; n < n_hi; ++ci) {
        const int w = chunk_width(ci, n, n_hi);
        const uint32_t ab = ci & 1, aph = (ci >> 1) & 1;
        const int nw0 = n + hf * 64;                 // first column of this warp's 64-column share of the chunk
        const bool has_cols = hf * 64 < w;           // warp-uniform
        if (has_cols && bias) {
          float bv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) bv[u] = (nw0 + u * 32 + lane < a.N) ? __ldg(bias + nw0 + u * 32 + lane) : 0.f;
          __syncwarp();
#pragma unroll
          for (int u = 0; u < 2; ++u) bias_s[u * 32 + lane] = bv[u];
          __syncwarp();
        }
        float4 rv[8];
        auto fetch_res = [&](int nc) {  // residual of the 32-column sub-chunk starting at column nc, coalesced layout
          const int nn = nc + xj * 4;
          const float* rp = a.res + boff + row0 * a.ldres + nn;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8, rp += 4 * a.ldres)
            rv[i8] = (row0 + 4 * i8 < a.M && nn < a.N) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if constexpr (kRes) { if (has_cols) fetch_res(nw0); }
        mbar_wait(&acc_full[ab], aph);
        tc_fence_after();
        const uint32_t taddr = tmem + (ab ? pg.acc1 : pg.acc0) + hf * 64 + ((uint32_t)(q * 32) << 16);
        if (!has_cols || (a.dbg & 32)) {
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
        } else if constexpr (kWide) {
          float v[64];
          tmem_ld32_issue(taddr, v);
          tmem_ld32_issue(taddr + 32, v + 32);
          tmem_wait_ld();
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
          if (pre != 1.f) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] *= pre;
          }
          if (bias) {
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + e);
              v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
            }
          }
          if (a.relu) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          if (a.row_post) {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] *= post;
          }
          const long ooff = boff + row0 * a.ldo + nw0 + xj * 8;
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
            bf16* op = (img == 0 ? a.out_hi : a.out_lo) + ooff;
#pragma unroll
            for (int i8 = 0; i8 < 8; ++i8, op += 4 * a.ldo) {
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
            if (pre != 1.f) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] *= pre;
            }
            if (bias) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + e);
                v[e] += bv.x; v[e + 1] += bv.y; v[e + 2] += bv.z; v[e + 3] += bv.w;
              }
            }
            if (a.relu) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (a.row_post) {
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] *= post;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(xp + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int nn = n0 + xj * 4;
            const bool col_ok = nn < a.N;
            const float* xr = xp + (lane >> 3) * 32;
            float* cp = kC ? a.C + boff + row0 * a.ldc + nn : nullptr;
            bf16* hp = kHi ? a.out_hi + boff + row0 * a.ldo + nn : nullptr;
            bf16* lp = kLo ? a.out_lo + boff + row0 * a.ldo + nn : nullptr;
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
              if constexpr (kC) cp += 4 * a.ldc;
              if constexpr (kHi) hp += 4 * a.ldo;
              if constexpr (kLo) lp += 4 * a.ldo;
            }
            __syncwarp();
            if constexpr (kRes) { if (c0 == 0 && hf * 64 + 32 < w) fetch_res(n0 + 32); }
          }
        }
        n += w;
        cnt += KBT;  // the ring stages that carried this chunk's weight blocks (the staging warps index the ring by cnt)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void split_bf16_kernel(const float* __restrict__ src, long ld, int rows, int cols, bf16* __restrict__ hi,
                                  bf16* __restrict__ lo) {
  pdl_sync();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n4 = (long)rows * (cols / 4);
  if (i >= n4) return;
  const int r = (int)(i / (cols / 4)), c = (int)(i % (cols / 4)) * 4;
  const float4 v = *reinterpret_cast<const float4*>(src + (long)r * ld + c);
  const float x[4] = {v.x, v.y, v.z, v.w};
  uint32_t h[2], l[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    h[u] = pack_bf16(x[2 * u], x[2 * u + 1]);
    l[u] = pack_bf16(x[2 * u] - bf16_round(x[2 * u]), x[2 * u + 1] - bf16_round(x[2 * u + 1]));
  }
  *reinterpret_cast<uint2*>(hi + (long)r * cols + c) = make_uint2(h[0], h[1]);
  if (lo) *reinterpret_cast<uint2*>(lo + (long)r * cols + c) = make_uint2(l[0], l[1]);
}

// C[m][n] = (sum_s part[s][m][n] + bias[n]) * row_post[m] + res[m][n]: the fixed-order reduction of a split-K GEMM's partial sums
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int S, long stride, const float* __restrict__ bias,
                                     const float* __restrict__ row_post, const float* res, long ldres, float* C, long ldc, int M, int N) {
  pdl_sync();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n4 = N / 4;
  if (i >= (long)M * n4) return;
  const int m = (int)(i / n4), n = (int)(i % n4) * 4;
  float4 acc = *reinterpret_cast<const float4*>(part + (long)m * N + n);
  for (int s = 1; s < S; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(part + s * stride + (long)m * N + n);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (row_post) {
    const float p = row_post[m];
    acc.x *= p; acc.y *= p; acc.z *= p; acc.w *= p;
  }
  if (res) {
    const float4 r = *reinterpret_cast<const float4*>(res + (long)m * ldres + n);
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
  }
  *reinterpret_cast<float4*>(C + (long)m * ldc + n) = acc;
}

}  // namespace

// C = (A W^T + bias) * row_post + res with a long reduction (K a few thousand) on FEW rows: the K range is cut into slices of 320,
// every (row panel, slice) pair is one work item of the panel kernel (A slice resident in tensor memory, TS-mode MMAs) writing
// its partial sum to `scratch` ([slices][M][N] fp32), and one small kernel adds the partials in slice order.  With M / 128 row
// panels well below the SM count the tile kernel would leave most SMs idle, each of the busy ones walking the whole K range.
void gemm_tc_splitk(const TcGemm& g0, float* scratch, cudaStream_t st) {
  S2S_CHECK(g0.nb * g0.nh == 1 && g0.passes == 3 && !g0.K2 && !g0.b_mn && !g0.out_hi && g0.C && g0.N % 4 == 0 && g0.ldc % 4 == 0 && g0.ldres % 4 == 0 &&
                !g0.relu && !g0.row_pre && g0.alpha == 1.f, "gemm_tc_splitk: unsupported variant");
  const int S = ceil_div(g0.K, 320);
  TcGemm g = g0;
  g.nh = S; g.a_ch = 320; g.b_ch = 320; g.K = 320;  // columns past K read as zeros (TMA out-of-bounds fill): the last slice may be partial
  g.bias = nullptr; g.row_post = nullptr; g.res = nullptr; g.ldres = 0;
  g.C = scratch; g.ldc = g0.N; g.sCb = 0; g.sCh = (long)g0.M * g0.N;
  g.force_panel = 1;
  gemm_tc(g, st);
  S2S_PROF("splitk_reduce", st);
  launch_pdl(splitk_reduce_kernel, ceil_div((long)g0.M * (g0.N / 4), 256), 256, 0, st, (const float*)scratch, S, (long)g0.M * g0.N, g0.bias, g0.row_post,
             g0.res, g0.ldres, g0.C, g0.ldc, g0.M, g0.N);
  S2S_LAUNCH_CHECK();
}

void split_bf16(const float* src, long ld, int rows, int cols, bf16* hi, bf16* lo, cudaStream_t st) {
  S2S_CHECK(cols % 4 == 0 && ld % 4 == 0, "split_bf16: width must be a multiple of 4");
  S2S_PROF("split_bf16", st);
  launch_pdl(split_bf16_kernel, ceil_div((long)rows * (cols / 4), 256), 256, 0, st, src, ld, rows, cols, hi, lo);
  S2S_LAUNCH_CHECK();
}

void gemm_tc(const TcGemm& g, cudaStream_t st) {
  S2S_CHECK(g.K % 16 == 0 && g.K > 0, "gemm_tc: K must be a positive multiple of 16");
  S2S_CHECK(g.passes == 1 || g.passes == 3, "gemm_tc: passes must be 1 or 3");
  S2S_CHECK(g.A_hi && g.B_hi && (g.passes == 1 || (g.A_lo && g.B_lo)), "gemm_tc: missing operand");
  const CUtensorMap mAh = make_bf16_2d_map(g.A_hi, g.a_rows, g.a_cols, g.a_pitch);
  S2S_CHECK(g.K2 == 0 || (g.passes == 1 && g.A2 && g.B2 && g.K2 % 16 == 0), "gemm_tc: a second K segment needs passes == 1 and both operands");
  const CUtensorMap mAl = g.passes == 3 ? make_bf16_2d_map(g.A_lo, g.a_rows, g.a_cols, g.a_pitch)
                          : g.K2     ? make_bf16_2d_map(g.A2, g.a2_rows, g.a2_cols, g.a2_pitch) : mAh;
  S2S_CHECK(!g.b_mn || g.K2 == 0, "gemm_tc: an MN-major B operand cannot be combined with a second K segment");
  S2S_CHECK(!g.b_mn || ((g.b_cb | g.b_ch) % 8 == 0), "gemm_tc: MN-major B boxes must start on 16-byte boundaries (column offsets % 8)");
  const int b_box = g.b_mn ? 64 : 128;
  const CUtensorMap mBh = make_bf16_2d_map(g.B_hi, g.b_rows, g.b_cols, g.b_pitch, b_box);
  const CUtensorMap mBl = g.passes == 3 ? make_bf16_2d_map(g.B_lo, g.b_rows, g.b_cols, g.b_pitch, b_box)
                          : g.K2     ? make_bf16_2d_map(g.B2, g.b2_rows, g.b2_cols, g.b2_pitch) : mBh;
  TcKernelArgs k;
  k.dbg = 0;
  k.K2 = g.K2; k.a2_cb = g.a2_cb; k.a2_ch = g.a2_ch; k.a2_rb = g.a2_rb; k.a2_rh = g.a2_rh;
  k.b2_cb = g.b2_cb; k.b2_ch = g.b2_ch; k.b2_rb = g.b2_rb; k.b2_rh = g.b2_rh;
  k.bias_sb = g.bias_sb; k.bias_sh = g.bias_sh;
  k.b_mn = g.b_mn;
  k.n_split = 1; k.n_per = 0;
  k.coalesced = g.N % 4 == 0 && g.ldc % 4 == 0 && g.ldres % 4 == 0 && g.ldo % 4 == 0 && g.sCb % 4 == 0 && g.sCh % 4 == 0 && !g.out_vt;
  if (const char* e = getenv("S2S_GEMM_COALESCED")) k.coalesced = k.coalesced && atoi(e) != 0;  // A/B timing only
  k.a_cb = g.a_cb; k.a_ch = g.a_ch; k.a_rb = g.a_rb; k.a_rh = g.a_rh;
  k.b_cb = g.b_cb; k.b_ch = g.b_ch; k.b_rb = g.b_rb; k.b_rh = g.b_rh;
  k.M = g.M; k.N = g.N; k.K = g.K; k.nb = g.nb; k.nh = g.nh; k.passes = g.passes; k.relu = g.relu; k.vt_L = g.vt_L;
  k.vt_col0 = g.vt_col0; k.vt_stride = g.vt_stride; k.vt_off = g.vt_off; k.vt_width = g.vt_width; k.vt_heads = g.vt_heads; k.out_vt_lo = g.out_vt_lo;
  k.alpha = g.alpha; k.bias = g.bias; k.row_pre = g.row_pre; k.row_post = g.row_post; k.res = g.res;
  k.C = g.C; k.ldc = g.ldc; k.sCb = g.sCb; k.sCh = g.sCh; k.ldres = g.ldres;
  k.out_hi = g.out_hi; k.out_lo = g.out_lo; k.ldo = g.ldo; k.out_vt = g.out_vt;
  if (const char* e = getenv("S2S_GEMM_DEBUG")) {  // timing experiments only
    const int d = atoi(e);
    if (d & 1) k.out_vt = nullptr;
    if (d & 2) k.out_hi = nullptr;
    if (d & 4) k.C = nullptr;
    k.dbg = d;
  }
  const int smem = G_SMEM;  // the kernel's extern array is declared __align__(1024) and checks it
  const int tiles = g.nb * g.nh * ceil_div(g.M, TM) * ceil_div(g.N, 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  S2S_PROF(g_profile_on ? prof_intern("gemm_tc M" + std::to_string(g.M) + " N" + std::to_string(g.N) +  " K" + std::to_string(g.K + g.K2) + " p" + std::to_string(g.passes) + " b" + std::to_string(g.nb * g.nh)) : "gemm_tc", st);
  int epi = -1;
  if (k.coalesced) {
    epi = (k.C ? 1 : 0) | (k.out_hi ? 2 : 0) | (k.out_lo ? 4 : 0) | (k.res ? 8 : 0);
    S2S_CHECK(!(epi & 4) || (epi & 2), "gemm_tc: a lo image needs a hi image");
    static const int wide = [] { const char* e = getenv("S2S_GEMM_WIDE"); return e ? atoi(e) : 1; }();  // A/B timing only
    if (wide && (epi == 2 || epi == 6) && g.N % 64 == 0 && g.ldo % 8 == 0 && g.sCb % 8 == 0 && g.sCh % 8 == 0) epi |= 16;
  }
  // panel kernel (A resident in tensor memory): non-batched, K a multiple of 64 that fits beside two accumulators
  static const int use_panel = [] { const char* e = getenv("S2S_GEMM_PANEL"); return e ? atoi(e) : 1; }();  // 0: A/B timing
  const int kbt = g.K / 64 + ceil_div(g.K2, 64);  // K blocks of the panel (both segments)
  const int a_cols = kbt * 32 * (g.passes == 3 ? 2 : 1);
  // Batched (decoy, head) GEMMs, a second K segment and an MN-major B are supported by the panel kernel but measured SLOWER there
  // (q.k^T+points 75 vs 60 us, P.v 60 vs 48, P.v_pts 48 vs 41 per launch at cfg2): with only two output chunks per panel the A
  // staging -> MMA -> epilogue chain of a panel is not overlapped with the next panel.  Off by default (S2S_GEMM_PANEL_BATCHED=1: A/B).
  static const int panel_batched = [] { const char* e = getenv("S2S_GEMM_PANEL_BATCHED"); return e ? atoi(e) : 0; }();
  const bool plain = g.nb * g.nh == 1 && g.K2 == 0 && !g.b_mn;
  if (use_panel && epi >= 0 && (plain || panel_batched || g.force_panel) && g.K % 64 == 0 && a_cols <= 320 && (g.K2 == 0 || g.passes == 1) &&
      (!g.b_mn || g.K2 == 0)) {
    PanelGeom pg;
    pg.acol_lo = kbt * 32;
    pg.acc0 = a_cols;
    pg.acc1 = a_cols + 128;
    pg.w1 = 512 - a_cols - 128 >= 128 ? 128 : 64;
    const int n_panels = g.nb * g.nh * ceil_div(g.M, TM);
    // fewer panels than half the SMs: split every panel's output columns over several CTAs, in units of 64 columns.  Measured at
    // L = 64 x 32 decoys (16 panels): the q|k|v projection (N = 6144) took 55 us on 16 SMs, each streaming the whole 3 MB weight
    // matrix (profiles/r02d_launch_summary_L64_B32_separate.txt); whole-step effect of the split with at least 128 / 64 columns
    // per CTA: 130 -> 183 -> 216 conformations/s at L = 64 x 32, 46.7 -> 74.4 -> 98.0 at cfg1 (L = 64, 1 decoy).
    k.n_split = 1; k.n_per = (g.N + 63) / 64 * 64;
    static const int split_env = [] { const char* e = getenv("S2S_GEMM_NSPLIT"); return e ? atoi(e) : 1; }();  // minimum 64-column units per CTA; 0: no split (A/B timing)
    if (split_env && n_panels * 2 <= sm_count()) {
      const int units = ceil_div(g.N, 64), want = std::min(sm_count() / n_panels, units / split_env);
      if (want > 1) {
        k.n_per = ceil_div(units, want) * 64;
        k.n_split = ceil_div(g.N, k.n_per);
      }
    }
    const int n_work = n_panels * k.n_split;
    const int pgrid = n_work < sm_count() ? n_work : sm_count();
    static bool pconf[24] = {};
    auto plaunch = [&](auto kern) {
      if (!pconf[epi]) {
        S2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        pconf[epi] = true;
      }
      launch_pdl(kern, pgrid, G_THREADS, smem, st, mAh, mAl, mBh, mBl, k, pg);
    };
    switch (epi) {
      case 1: plaunch(gemm_panel_kernel<1>); break;
      case 2: plaunch(gemm_panel_kernel<2>); break;
      case 3: plaunch(gemm_panel_kernel<3>); break;
      case 6: plaunch(gemm_panel_kernel<6>); break;
      case 7: plaunch(gemm_panel_kernel<7>); break;
      case 9: plaunch(gemm_panel_kernel<9>); break;
      case 11: plaunch(gemm_panel_kernel<11>); break;
      case 15: plaunch(gemm_panel_kernel<15>); break;
      case 18: plaunch(gemm_panel_kernel<18>); break;
      case 22: plaunch(gemm_panel_kernel<22>); break;
      default: S2S_CHECK(false, "gemm_tc: unexpected epilogue variant");
    }
    S2S_LAUNCH_CHECK();
    return;
  }
  static bool configured[24] = {};  // per instantiation (index epi + 1): every gemm_tc_kernel<EPI> has the same pointer type
  auto launch = [&](auto kern) {
    if (!configured[epi + 1]) {
      S2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      configured[epi + 1] = true;
    }
    launch_pdl(kern, grid, G_THREADS, smem, st, mAh, mAl, mBh, mBl, k);
  };
  switch (epi) {
    case 1: launch(gemm_tc_kernel<1>); break;
    case 2: launch(gemm_tc_kernel<2>); break;
    case 3: launch(gemm_tc_kernel<3>); break;
    case 6: launch(gemm_tc_kernel<6>); break;
    case 7: launch(gemm_tc_kernel<7>); break;
    case 9: launch(gemm_tc_kernel<9>); break;
    case 11: launch(gemm_tc_kernel<11>); break;
    case 15: launch(gemm_tc_kernel<15>); break;
    case 18: launch(gemm_tc_kernel<18>); break;
    case 22: launch(gemm_tc_kernel<22>); break;
    default: epi = -1; launch(gemm_tc_kernel<-1>); break;
  }
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
