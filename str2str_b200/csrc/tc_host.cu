// Host-side helpers shared by the tcgen05 kernels: TMA tensor maps, SM count, pre-swizzled weight blocks.
#include "s2s_internal.cuh"
#include "tc_common.cuh"

namespace s2s {

using namespace tc;

namespace {

// one [128 x 64] bf16 weight block src[n0 .. n0+128)[k0 .. k0+64) in the SW128 K-major layout
__global__ void build_wtile_kernel(const float* __restrict__ src, int ld, int n0, int k0, unsigned char* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TM * KBLK) return;
  const int r = idx / KBLK, c = idx % KBLK;
  *reinterpret_cast<bf16*>(dst + sw128_offset(r, c)) = __float2bfloat16_rn(src[(size_t)(n0 + r) * ld + k0 + c]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  pdl_sync();
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

// [rows][cols] bf16 row-major tensor (row pitch in elements), boxes of box_rows rows x 64 columns, 128-byte swizzle (or, with
// swizzle = false, a verbatim copy: used for pre-swizzled weight images); out-of-bounds box elements read as zero
CUtensorMap make_bf16_2d_map(const void* base, size_t rows, size_t cols, size_t row_pitch_elems, int box_rows, bool swizzle) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult rc = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

size_t ee_wimg_elems() { return (size_t)4 * TM * KBLK; }

// edge-embedder weight image: W2 k-blocks 0, 1, then W3 k-blocks 0, 1 (the order the pipelined kernel keeps them resident)
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
  launch_pdl(f32_to_bf16_kernel, ceil_div(n, 256), 256, 0, st, src, dst, n);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
