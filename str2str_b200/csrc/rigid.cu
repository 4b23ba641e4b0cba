// Rigid-frame and SE(3) diffusion kernels: frame update, IGSO(3)/VP-SDE score + reverse step,
// forward perturbation, idealised backbone atoms.  One thread per residue; one CTA per decoy where a
// per-decoy reduction (centre of mass) is needed.  dtype flow follows the reference exactly: fp32 where the
// reference is fp32, double where its fp64 masks promote the arithmetic.
#include "s2s_internal.cuh"

namespace s2s {

namespace {

template <typename T> struct M;
template <> struct M<float> {
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
};
template <> struct M<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double sin_(double x) { return sin(x); }
  static __device__ __forceinline__ double cos_(double x) { return cos(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
};

// rotation3d.matrix_to_quaternion :102-161 — candidate with the largest |component| (first index on ties)
template <typename T>
__device__ void mat_to_quat(const T m[9], T q[4]) {
  const T m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
  T s[4] = {T(1) + m00 + m11 + m22, T(1) + m00 - m11 - m22, T(1) - m00 + m11 - m22, T(1) - m00 - m11 + m22};
  T qa[4];
  int pick = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    qa[k] = s[k] > T(0) ? M<T>::sqrt_(s[k]) : T(0);
    if (qa[k] > qa[pick]) pick = k;
  }
  T c[4];
  if (pick == 0) { c[0] = qa[0] * qa[0]; c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
  else if (pick == 1) { c[0] = m21 - m12; c[1] = qa[1] * qa[1]; c[2] = m10 + m01; c[3] = m02 + m20; }
  else if (pick == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = qa[2] * qa[2]; c[3] = m12 + m21; }
  else { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = qa[3] * qa[3]; }
  const T den = T(2) * (qa[pick] > T(0.1) ? qa[pick] : T(0.1));
#pragma unroll
  for (int k = 0; k < 4; ++k) q[k] = c[k] / den;
}

template <typename T>
__device__ __forceinline__ T sin_half_over_angle(T ang, T half) {
  const T aa = ang < T(0) ? -ang : ang;
  return aa < T(1e-6) ? T(0.5) - ang * ang / T(48) : M<T>::sin_(half) / ang;
}
// rotation3d.quaternion_to_axis_angle :525-553 (no sign standardisation)
template <typename T>
__device__ void quat_to_aa(const T q[4], T v[3]) {
  const T n = M<T>::sqrt_(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T half = M<T>::atan2_(n, q[0]);
  const T ang = T(2) * half;
  const T s = sin_half_over_angle(ang, half);
  v[0] = q[1] / s; v[1] = q[2] / s; v[2] = q[3] / s;
}
// rotation3d.axis_angle_to_quaternion :493-522
template <typename T>
__device__ void aa_to_quat(const T v[3], T q[4]) {
  const T ang = M<T>::sqrt_(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const T half = ang * T(0.5);
  const T s = sin_half_over_angle(ang, half);
  q[0] = M<T>::cos_(half); q[1] = v[0] * s; q[2] = v[1] * s; q[3] = v[2] * s;
}
// rotation3d.quaternion_to_matrix :41-70 (normalising form)
template <typename T>
__device__ void quat_to_mat_norm(const T q[4], T R[9]) {
  const T r = q[0], i = q[1], j = q[2], k = q[3];
  const T s = T(2) / (r * r + i * i + j * j + k * k);
  R[0] = T(1) - s * (j * j + k * k); R[1] = s * (i * j - k * r); R[2] = s * (i * k + j * r);
  R[3] = s * (i * j + k * r); R[4] = T(1) - s * (i * i + k * k); R[5] = s * (j * k - i * r);
  R[6] = s * (i * k - j * r); R[7] = s * (j * k + i * r); R[8] = T(1) - s * (i * i + j * j);
}
template <typename T>
__device__ void aa_to_mat(const T v[3], T R[9]) {
  T q[4];
  aa_to_quat(v, q);
  quat_to_mat_norm(q, R);
}
template <typename T>
__device__ void mat_to_aa(const T R[9], T v[3]) {
  T q[4];
  mat_to_quat(R, q);
  quat_to_aa(q, v);
}
__device__ __forceinline__ void quat_mul(const float p[4], const float q[4], float o[4]) {
  o[0] = p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3];
  o[1] = p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2];
  o[2] = p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1];
  o[3] = p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0];
}
// so3.compose_rotvec :13-19 — fp64 product of the two rotations; first operand is fp32, second is T2
template <typename T2>
__device__ void compose_rotvec(const float v1[3], const T2 v2[3], float out[3]) {
  float R1f[9];
  aa_to_mat<float>(v1, R1f);
  T2 R2t[9];
  aa_to_mat<T2>(v2, R2t);
  double C[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      C[r * 3 + c] = (double)R1f[r * 3] * (double)R2t[c] + (double)R1f[r * 3 + 1] * (double)R2t[3 + c] +
                     (double)R1f[r * 3 + 2] * (double)R2t[6 + c];
  double v[3];
  mat_to_aa<double>(C, v);
  out[0] = (float)v[0]; out[1] = (float)v[1]; out[2] = (float)v[2];
}

// d/domega log IGSO3 density: the reference's fp32 series (so3.py:58,121-125,130), ascending l, per-term
// division, no fused multiply-adds.  Terms with exp(-l(l+1)sigma^2/2) below ~1e-26 cannot change either fp32
// sum, so the loop stops there instead of at l = 999.
__device__ float igso3_score_scale(float omega, float sigma) {
  const float lo = sinf(__fmul_rn(omega, 0.5f));      // sin(omega/2)
  const float dlo = __fmul_rn(0.5f, cosf(__fmul_rn(omega, 0.5f)));
  const float lo2 = __fmul_rn(lo, lo);
  const float s2h = __fdiv_rn(__fmul_rn(sigma, sigma), 2.0f);  // eps**2 / 2
  float f = 0.f, df = 0.f;
  for (int l = 0; l < 1000; ++l) {
    const float x = __fmul_rn((float)(-(l * (l + 1))), s2h);
    if (x < -60.f) break;
    const float e = expf(x);
    const float c = __fmul_rn((float)(2 * l + 1), e);
    const float lh = (float)l + 0.5f;
    const float arg = __fmul_rn(omega, lh);
    const float hi = sinf(arg), ch = cosf(arg);
    f = __fadd_rn(f, __fdiv_rn(__fmul_rn(c, hi), lo));
    const float dhi = __fmul_rn(lh, ch);
    const float num = __fadd_rn(__fmul_rn(lo, dhi), -__fmul_rn(hi, dlo));
    df = __fadd_rn(df, __fdiv_rn(__fmul_rn(c, num), lo2));
  }
  return __fdiv_rn(df, __fadd_rn(f, 1e-4f));
}

// ---- frame update (rigid_utils.py:1042-1066, 590-619) -------------------------------------------------
__global__ void frame_update_kernel(float* __restrict__ quat, float* __restrict__ trans,
                                    const float* __restrict__ upd, const float* __restrict__ diffuse, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float q[4] = {quat[r * 4], quat[r * 4 + 1], quat[r * 4 + 2], quat[r * 4 + 3]};
  const float m = diffuse[r];
  const float* u = upd + (long)r * 6;
  float R[9];
  quat_to_rot(q, R);
  const float vq[4] = {0.f, u[0], u[1], u[2]};
  float dq[4];
  quat_mul(q, vq, dq);
  float nq[4];
  float n2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nq[k] = q[k] + dq[k] * m;
    n2 += nq[k] * nq[k];
  }
  const float n = sqrtf(n2);
#pragma unroll
  for (int k = 0; k < 4; ++k) quat[r * 4 + k] = nq[k] / n;
  trans[r * 3 + 0] += (R[0] * u[3] + R[1] * u[4] + R[2] * u[5]) * m;
  trans[r * 3 + 1] += (R[3] * u[3] + R[4] * u[4] + R[5] * u[5]) * m;
  trans[r * 3 + 2] += (R[6] * u[3] + R[7] * u[4] + R[8] * u[5]) * m;
}

// BackboneUpdate + compose_q_update_vec in one pass (layers.py:232-241, rigid_utils.py:1042-1066): one warp per residue
// computes the six outputs of Linear(256 -> 6)(node * diffuse_mask) in exact fp32 and applies them to the frame.
__global__ void __launch_bounds__(256) bb_update_frame_kernel(const float* __restrict__ node, const float* __restrict__ W,
                                                              const float* __restrict__ bias, float* __restrict__ quat,
                                                              float* __restrict__ trans, const float* __restrict__ diffuse,
                                                              int rows) {
  pdl_sync();
  const int lane = threadIdx.x % 32;
  const int r = blockIdx.x * 8 + threadIdx.x / 32;
  if (r >= rows) return;
  const float4 x0 = *reinterpret_cast<const float4*>(node + (long)r * C_S + lane * 8);
  const float4 x1 = *reinterpret_cast<const float4*>(node + (long)r * C_S + lane * 8 + 4);
  float u[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + k * C_S + lane * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + k * C_S + lane * 8 + 4));
    float acc = x0.x * w0.x;
    acc = fmaf(x0.y, w0.y, acc); acc = fmaf(x0.z, w0.z, acc); acc = fmaf(x0.w, w0.w, acc);
    acc = fmaf(x1.x, w1.x, acc); acc = fmaf(x1.y, w1.y, acc); acc = fmaf(x1.z, w1.z, acc); acc = fmaf(x1.w, w1.w, acc);
    u[k] = warp_sum(acc);
  }
  if (lane != 0) return;
  const float m = diffuse[r];
#pragma unroll
  for (int k = 0; k < 6; ++k) u[k] = u[k] * m + bias[k];
  float q[4] = {quat[r * 4], quat[r * 4 + 1], quat[r * 4 + 2], quat[r * 4 + 3]};
  float R[9];
  quat_to_rot(q, R);
  const float vq[4] = {0.f, u[0], u[1], u[2]};
  float dq[4];
  quat_mul(q, vq, dq);
  float nq[4];
  float n2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nq[k] = q[k] + dq[k] * m;
    n2 += nq[k] * nq[k];
  }
  const float n = sqrtf(n2);
#pragma unroll
  for (int k = 0; k < 4; ++k) quat[r * 4 + k] = nq[k] / n;
  trans[r * 3 + 0] += (R[0] * u[3] + R[1] * u[4] + R[2] * u[5]) * m;
  trans[r * 3 + 1] += (R[3] * u[3] + R[4] * u[4] + R[5] * u[5]) * m;
  trans[r * 3 + 2] += (R[6] * u[3] + R[7] * u[4] + R[8] * u[5]) * m;
}

__global__ void split_rigids_kernel(const float* __restrict__ rig, float* __restrict__ quat,
                                    float* __restrict__ trans, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) quat[r * 4 + k] = rig[(long)r * 7 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) trans[r * 3 + k] = rig[(long)r * 7 + 4 + k] * 0.1f;  // Angstrom -> nm (ipa.py:339)
}
__global__ void join_rigids_kernel(const float* __restrict__ quat, const float* __restrict__ trans,
                                   float* __restrict__ rig, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) rig[(long)r * 7 + k] = quat[r * 4 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) rig[(long)r * 7 + 4 + k] = __fdiv_rn(trans[r * 3 + k], 0.1f);  // ipa.py:379
}

// ---- score + reverse (frame.py:109-210, so3.py:274-371, r3.py:79-137) ---------------------------------
__global__ void __launch_bounds__(256) se3_step_kernel(Se3StepArgs a) {
  pdl_sync();
  extern __shared__ double xs[];  // [L][3] un-centred new translations (nm)
  __shared__ double red[3][8];
  const int b = blockIdx.x, L = a.L, tid = threadIdx.x;
  const float* sf = a.sched_f + b * 8;
  const float sigma_q = sf[1], g_rot = sf[2], g2_rot = sf[3], e_half = sf[4], cvar = sf[5], b_t = sf[6], g_tr = sf[7];
  const double dt = a.sched_d[b * 2], sqrt_dt = a.sched_d[b * 2 + 1];
  const double half = a.probability_flow ? 0.5 : 1.0;
  double acc[3] = {0.0, 0.0, 0.0};

  for (int l = tid; l < L; l += 256) {
    const long r = (long)b * L + l;
    const float* rt = a.rig_t + r * 7;
    float qt[4] = {rt[0], rt[1], rt[2], rt[3]};
    float Rt[9];
    quat_to_rot(qt, Rt);
    double rs[3], ts[3];
    if (a.mode != 2) {
      const float* r0 = a.rig_0 + r * 7;
      // frame.py:119-127: quaternion of R0^-1 (through its rotation matrix), times quaternion of R_t
      const float n2 = r0[0] * r0[0] + r0[1] * r0[1] + r0[2] * r0[2] + r0[3] * r0[3];
      float q0i[4] = {r0[0] / n2, -r0[1] / n2, -r0[2] / n2, -r0[3] / n2};
      float R0i[9], qa[4], qb[4], q0t[4], vec[3];
      quat_to_rot(q0i, R0i);
      mat_to_quat<float>(R0i, qa);
      mat_to_quat<float>(Rt, qb);
      quat_mul(qa, qb, q0t);
      quat_to_aa<float>(q0t, vec);
      const float omega = sqrtf(vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2]) + 1e-6f;
      const float sc = igso3_score_scale(omega, sigma_q);
      const float den = omega + 1e-6f;
      const double mk = (double)a.mask[r];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rs[k] = (double)__fdiv_rn(__fmul_rn(sc, vec[k]), den) * mk;
        const float xt = __fmul_rn(rt[4 + k], 0.1f), x0 = __fmul_rn(r0[4 + k], 0.1f);
        ts[k] = (double)(-__fdiv_rn(__fadd_rn(xt, -__fmul_rn(e_half, x0)), cvar)) * mk;
      }
      if (a.rot_score) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          a.rot_score[r * 3 + k] = rs[k];
          a.trans_score[r * 3 + k] = ts[k];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rs[k] = a.rot_score[r * 3 + k];
        ts[k] = a.trans_score[r * 3 + k];
      }
    }
    if (a.mode == 1) continue;

    // rotation: rot_{t-1} = rot_t o Exp(-perturb)   (so3.py:357-370)
    float rotvec_t[3];
    mat_to_aa<float>(Rt, rotvec_t);
    double neg_p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double p = (double)(-1.0f * g2_rot) * rs[k] * dt * half;
      if (!a.probability_flow) p += (double)g_rot * sqrt_dt * ((double)a.noise_scale * (double)a.rot_noise[r * 3 + k]);
      neg_p[k] = -1.0 * p;
    }
    float rotvec_n[3];
    compose_rotvec<double>(rotvec_t, neg_p, rotvec_n);
    // translation drift in nm (r3.py:101-116)
    double xn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float x = __fmul_rn(rt[4 + k], 0.1f);
      const float f_t = __fmul_rn(__fmul_rn(-0.5f, b_t), x);
      double p = ((double)f_t - (double)__fmul_rn(g_tr, g_tr) * ts[k]) * dt * half;
      if (!a.probability_flow) p += (double)g_tr * sqrt_dt * ((double)a.noise_scale * (double)a.trans_noise[r * 3 + k]);
      xn[k] = (double)x - p;
      xs[l * 3 + k] = xn[k];
      acc[k] += xn[k];
    }
    // new rotation (masked), re-encoded as a quaternion (frame.py:206-210,9-15; rigid_utils.py:1203-1215)
    const double m = a.diffuse ? (double)a.diffuse[r] : 1.0;
    double rv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rv[k] = m * (double)rotvec_n[k] + (1.0 - m) * (double)rotvec_t[k];
    double Rn[9];
    aa_to_mat<double>(rv, Rn);
    float Rnf[9], qn[4];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rnf[k] = (float)Rn[k];
    mat_to_quat<float>(Rnf, qn);
    float* o = a.rig_out + r * 7;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = qn[k];
  }
  if (a.mode == 1) return;

  // centre of mass over ALL L rows (mask=None in the reference call, r3.py:117-122; frame.py:195-203)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[k][tid >> 5] = v;
  }
  __syncthreads();
  double com[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[k][w];
    com[k] = v / (double)(float)L;
  }
  for (int l = tid; l < L; l += 256) {
    const long r = (long)b * L + l;
    const double m = a.diffuse ? (double)a.diffuse[r] : 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double xnew = (xs[l * 3 + k] - com[k]) / 0.1;  // unscale: fp64 tensor / python float 0.1
      a.rig_out[r * 7 + 4 + k] = (float)(m * xnew + (1.0 - m) * (double)a.rig_t[r * 7 + 4 + k]);
    }
  }
}

// ---- forward perturbation (frame.py:36-107, so3.py:244-272,315-331, r3.py:49-74) ----------------------
__global__ void se3_perturb_kernel(Se3PerturbArgs a) {
  pdl_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * a.L) return;
  const int b = idx / a.L;
  const long r = idx;
  float R0[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R0[k] = a.rot0[r * 9 + k];
  float rot0[3];
  mat_to_aa<float>(R0, rot0);
  // uniform axis, IGSO(3) angle by inverse-CDF interpolation (np.interp semantics, fp64)
  const float ax = a.axis_noise[r * 3], ay = a.axis_noise[r * 3 + 1], az = a.axis_noise[r * 3 + 2];
  const float an = sqrtf(ax * ax + ay * ay + az * az);
  const double* cdf = a.cdf + (long)b * 1000;
  const double u = (double)a.u_noise[r];
  double om;
  if (u <= cdf[0]) om = (double)a.omega_grid[0];
  else if (u >= cdf[999]) om = (double)a.omega_grid[999];
  else {
    int lo = 0, hi = 999;  // invariant: cdf[lo] <= u < cdf[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid; else hi = mid;
    }
    const double slope = ((double)a.omega_grid[lo + 1] - (double)a.omega_grid[lo]) / (cdf[lo + 1] - cdf[lo]);
    om = slope * (u - cdf[lo]) + (double)a.omega_grid[lo];
  }
  const float omf = (float)om;
  const float d[3] = {(ax / an) * omf, (ay / an) * omf, (az / an) * omf};
  float rot_t[3];
  compose_rotvec<float>(rot0, d, rot_t);
  const float e_half = a.sched_f[b * 2], sd = a.sched_f[b * 2 + 1];
  const float m = a.diffuse ? a.diffuse[r] : 1.f;
  float rv[3], xo[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float x0 = a.trans0[r * 3 + k];
    const float xt = __fdiv_rn(__fadd_rn(__fmul_rn(a.trans_noise[r * 3 + k], sd), __fmul_rn(e_half, __fmul_rn(x0, 0.1f))), 0.1f);
    xo[k] = m * xt + (1.f - m) * x0;
    rv[k] = m * rot_t[k] + (1.f - m) * rot0[k];
  }
  float Rn[9], qn[4];
  aa_to_mat<float>(rv, Rn);
  mat_to_quat<float>(Rn, qn);
  float* o = a.rig_out + r * 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = qn[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) o[4 + k] = xo[k];
}

// ---- idealised backbone (all_atom.py:141-173) ----------------------------------------------------------
// table rows (per residue type, 21 rows x 33 floats): pos[5][3] (N CA C O CB), mask[5], psi frame rot[9],
// psi frame trans[3], bb_valid
__global__ void backbone_atoms_kernel(const float* __restrict__ rig, const float* __restrict__ psi,
                                      const long long* __restrict__ aatype, const float* __restrict__ table,
                                      float* __restrict__ atom37, float* __restrict__ atom14, int rows) {
  pdl_sync();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int aa = aatype ? (int)aatype[r] : 0;
  const float* T = table + aa * 33;
  float q[4] = {rig[(long)r * 7], rig[(long)r * 7 + 1], rig[(long)r * 7 + 2], rig[(long)r * 7 + 3]};
  float R[9];
  quat_to_rot(q, R);
  const float tx = rig[(long)r * 7 + 4], ty = rig[(long)r * 7 + 5], tz = rig[(long)r * 7 + 6];
  const float s = psi[2 * r], c = psi[2 * r + 1];
  const float* D = T + 20;  // psi default rotation
  // Rpsi = D * Rx(psi),  Rx = [[1,0,0],[0,c,-s],[0,s,c]]   (all_atom.py:43-59)
  float Rp[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rp[i * 3] = D[i * 3];
    Rp[i * 3 + 1] = D[i * 3 + 1] * c + D[i * 3 + 2] * s;
    Rp[i * 3 + 2] = -D[i * 3 + 1] * s + D[i * 3 + 2] * c;
  }
  float local[5][3];
  const float valid = T[32];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k == 3) {
      const float* p = T + 9;
#pragma unroll
      for (int i = 0; i < 3; ++i) local[3][i] = Rp[i * 3] * p[0] + Rp[i * 3 + 1] * p[1] + Rp[i * 3 + 2] * p[2] + T[29 + i];
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) local[k][i] = T[k * 3 + i] * valid;
    }
  }
  float* a37 = atom37 + (long)r * 37 * 3;
  float* a14 = atom14 ? atom14 + (long)r * 14 * 3 : nullptr;
  for (int k = 0; k < 37 * 3; ++k) a37[k] = 0.f;
  if (a14)
    for (int k = 0; k < 14 * 3; ++k) a14[k] = 0.f;
  const int slot37[5] = {0, 1, 2, 4, 3};  // atom14 N,CA,C,O,CB -> atom37 N,CA,C,CB,O
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float mk = T[15 + k];
    const float gx = (R[0] * local[k][0] + R[1] * local[k][1] + R[2] * local[k][2] + tx) * mk;
    const float gy = (R[3] * local[k][0] + R[4] * local[k][1] + R[5] * local[k][2] + ty) * mk;
    const float gz = (R[6] * local[k][0] + R[7] * local[k][1] + R[8] * local[k][2] + tz) * mk;
    a37[slot37[k] * 3] = gx; a37[slot37[k] * 3 + 1] = gy; a37[slot37[k] * 3 + 2] = gz;
    if (a14) { a14[k * 3] = gx; a14[k * 3 + 1] = gy; a14[k * 3 + 2] = gz; }
  }
}

}  // namespace

void frame_update(float* quat, float* trans, const float* upd6, const float* diffuse, int rows, cudaStream_t st) {
  launch_pdl(frame_update_kernel, ceil_div(rows, 128), 128, 0, st, quat, trans, upd6, diffuse, rows);
  S2S_LAUNCH_CHECK();
}
void bb_update_frame(const float* node, const float* W, const float* bias, float* quat, float* trans, const float* diffuse,
                     int rows, cudaStream_t st) {
  S2S_PROF("bb_update_frame", st);
  launch_pdl(bb_update_frame_kernel, ceil_div(rows, 8), 256, 0, st, node, W, bias, quat, trans, diffuse, rows);
  S2S_LAUNCH_CHECK();
}
void split_rigids(const float* rig7, float* quat, float* trans_nm, int rows, cudaStream_t st) {
  launch_pdl(split_rigids_kernel, ceil_div(rows, 128), 128, 0, st, rig7, quat, trans_nm, rows);
  S2S_LAUNCH_CHECK();
}
void join_rigids(const float* quat, const float* trans_nm, float* rig7, int rows, cudaStream_t st) {
  launch_pdl(join_rigids_kernel, ceil_div(rows, 128), 128, 0, st, quat, trans_nm, rig7, rows);
  S2S_LAUNCH_CHECK();
}
void se3_step(const Se3StepArgs& a, cudaStream_t st) {
  S2S_CHECK(a.mode >= 0 && a.mode <= 2, "se3_step: bad mode");
  S2S_CHECK(a.probability_flow || (a.rot_noise && a.trans_noise) || a.mode == 1, "se3_step: SDE mode needs noise");
  const size_t smem = (size_t)a.L * 3 * sizeof(double);
  S2S_CHECK(smem <= 48 * 1024, "se3_step: chain too long (L <= 2048)");
  launch_pdl(se3_step_kernel, a.B, 256, smem, st, a);
  S2S_LAUNCH_CHECK();
}
void se3_perturb(const Se3PerturbArgs& a, cudaStream_t st) {
  launch_pdl(se3_perturb_kernel, ceil_div((long)a.B * a.L, 128), 128, 0, st, a);
  S2S_LAUNCH_CHECK();
}
void backbone_atoms(const float* rig7, const float* psi, const long long* aatype, const float* table,
                    float* atom37, float* atom14, int rows, cudaStream_t st) {
  launch_pdl(backbone_atoms_kernel, ceil_div(rows, 128), 128, 0, st, rig7, psi, aatype, table, atom37, atom14, rows);
  S2S_LAUNCH_CHECK();
}

}  // namespace s2s
