// Shared device/host helpers for the Str2Str B200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>
#include <utility>

namespace s2s {

// ---- model constants (reference configs/model/diffusion.yaml:20-40) ---------------------------------
constexpr int C_S = 256;      // node channels
constexpr int C_Z = 128;      // pair channels
constexpr int C_H = 256;      // IPA hidden per head
constexpr int N_H = 8;        // IPA heads
constexpr int P_Q = 8;        // query/key points
constexpr int P_V = 12;       // value points
constexpr int N_BLK = 4;
constexpr int D_SKIP = 64;
constexpr int D_TFM = 320;
constexpr int TFM_H = 4;
constexpr int TFM_HD = 80;
constexpr int D_ET = 384;     // EdgeTransition hidden = c_z + 2*(c_s/2)
constexpr int IPA_FEAT = N_H * (C_H + 4 * P_V + C_Z / 4);  // 2688
constexpr int N_BINS = 22;

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define S2S_CHECK(cond, msg)                                                             \
  do {                                                                                   \
    if (!(cond)) throw ::s2s::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)

#define S2S_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::s2s::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + #expr + ": " + cudaGetErrorString(_e)); \
  } while (0)

extern long long g_launch_count;  // kernels launched by this library (api.cu)
#define S2S_LAUNCH_CHECK()            \
  do {                                \
    ++::s2s::g_launch_count;          \
    S2S_CUDA(cudaGetLastError());     \
  } while (0)

// Optional per-kernel timing (bench.py roofline): when enabled, CUDA events bracket the named launches on the
// launching stream.  Off by default and never enabled while a CUDA graph is being captured.
struct ProfScope {
  const char* name;
  cudaStream_t st;
  cudaEvent_t e0 = nullptr;
  ProfScope(const char* n, cudaStream_t s);
  ~ProfScope();
};
extern bool g_profile_on;
const char* prof_intern(const std::string& name);  // stable storage for dynamically built profile names
#define S2S_PROF(name, st) ::s2s::ProfScope _prof_scope(name, st)

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch.  Every kernel of the denoising iteration is launched with the programmatic-stream-
// serialization attribute and calls pdl_sync() before its first access to global memory: the grid may then become resident
// (barrier init, tensor-memory allocation, index arithmetic) while the previous kernel of the stream drains, instead of
// paying launch latency + prologue after it.  pdl_sync() waits until the prerequisite grid has completed and its memory
// operations are visible, and only THEN allows the next launch: at most one successor is staged behind a running kernel.
// The attribute and the wait go together: a kernel launched through launch_pdl() without a pdl_sync() would race.
extern bool g_pdl;  // api.cu; S2S_PDL=0 launches every kernel fully serialised (A/B timing)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  S2S_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

typedef __nv_bfloat16 bf16;

// ---- small device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// warp-level bf16 MMA (legacy tensor path; used only for the small in-kernel GEMMs of the IPA pair kernel)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}

// quaternion (w,x,y,z) -> rotation matrix, quadratic form WITHOUT normalisation
// (reference src/common/rigid_utils.py:187-207)
__device__ __forceinline__ void quat_to_rot(const float q[4], float R[9]) {
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  R[0] = a * a + b * b - c * c - d * d;
  R[1] = 2.f * (b * c - a * d);
  R[2] = 2.f * (b * d + a * c);
  R[3] = 2.f * (b * c + a * d);
  R[4] = a * a - b * b + c * c - d * d;
  R[5] = 2.f * (c * d - a * b);
  R[6] = 2.f * (b * d - a * c);
  R[7] = 2.f * (c * d + a * b);
  R[8] = a * a - b * b - c * c + d * d;
}

}  // namespace s2s
