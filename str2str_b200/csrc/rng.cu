// Counter-based random draws for the sampler ("per-rank RNG = Philox(seed, subsequence = global decoy id) so
// results are independent of the world size").  The reference draws its perturbation / SDE noise from torch's global
// generators (so3.py:259-262, r3.py:66,109), whose streams depend on the batch a decoy happens to share and on the device;
// here decoy d of a job always sees Philox4x32-10 subsequence d of the job's seed, whichever rank and batch it lands in.
// Layout of a decoy's stream: draw `stream_id` (0 = perturbation axis, 1 = perturbation angle quantile, 2 = perturbation
// translation, 16 + 2k / 17 + 2k = rotation / translation noise of SDE iteration k) starts at offset stream_id * 2^24; element
// e of the draw is output e of that block (4 outputs per Philox counter).
#include <curand_kernel.h>

#include "s2s_internal.cuh"

namespace s2s {

namespace {

// decoy_ids / stream_ids (optional, per row): row b draws for decoy decoy_ids[b] (else first_decoy + b), block stream_ids[b]
// (else stream_id); rows with a negative decoy id are left untouched (idle rows of a continuous batch)
__global__ void philox_fill_kernel(float* __restrict__ out, int B, long n_per_decoy, unsigned long long seed, long long first_decoy,
                                   unsigned long long stream_id, int uniform, const long long* __restrict__ decoy_ids,
                                   const int* __restrict__ stream_ids) {
  pdl_sync();
  const long quads = (n_per_decoy + 3) / 4;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * quads) return;
  const int b = (int)(idx / quads);
  const long qd = idx - (long)b * quads;
  const long long decoy = decoy_ids ? decoy_ids[b] : first_decoy + b;
  if (decoy < 0) return;
  if (stream_ids) stream_id = (unsigned long long)stream_ids[b];
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)decoy, (stream_id << 24) + 4ull * (unsigned long long)qd, &st);
  float4 v;
  if (uniform) {
    v = curand_uniform4(&st);  // (0, 1]
    v.x = 1.f - v.x; v.y = 1.f - v.y; v.z = 1.f - v.z; v.w = 1.f - v.w;  // [0, 1) like torch.rand (so3.py:262)
  } else {
    v = curand_normal4(&st);
  }
  float* o = out + (long)b * n_per_decoy + 4 * qd;
  const long left = n_per_decoy - 4 * qd;
  o[0] = v.x;
  if (left > 1) o[1] = v.y;
  if (left > 2) o[2] = v.z;
  if (left > 3) o[3] = v.w;
}

}  // namespace

void philox_fill(float* out, int B, long n_per_decoy, unsigned long long seed, long long first_decoy, unsigned long long stream_id,
                 int uniform, cudaStream_t st, const long long* decoy_ids, const int* stream_ids) {
  S2S_CHECK(n_per_decoy > 0 && n_per_decoy < (1l << 24), "philox_fill: at most 2^24 - 1 elements per decoy and draw");
  const long quads = (n_per_decoy + 3) / 4;
  launch_pdl(philox_fill_kernel, ceil_div((long)B * quads, 256), 256, 0, st, out, B, n_per_decoy, seed, first_decoy, stream_id, uniform, decoy_ids,
                                                                     stream_ids);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
