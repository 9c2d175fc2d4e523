// normals.cu — K3: per-point 3x3 covariance + closed-form smallest eigenpair.
//
// Replaces pcl::NormalEstimationOMP::compute as called at main.cpp:104-109
// (computePointNormal -> computeMeanAndCovarianceMatrix -> solvePlaneParameters ->
// eigen33 -> computeRoots / computeRoots2, then flipNormalTowardsViewpoint with the
// default viewpoint (0,0,0)).  Scalar FP32, no tensor cores: this is k gathers and a 3x3
// eigenproblem per point, not a dense contraction.
//
// Bit parity with the oracle (oracle/iftwt_oracle.cpp: point_normal): the FP32
// operation order below is the oracle's, the build uses -fmad=false so nothing is
// contracted, division and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt), and
// the three elementary functions of the eigen solver are evaluated by the fixed
// binary64 series of DESIGN.md ("spec trig") and rounded once to binary32 — every
// step is a correctly rounded IEEE operation on both machines.
#include <stdlib.h>

#include "common.cuh"

#define SID_NONE 0xFFFFFFFFu

namespace {

__device__ __forceinline__ double spec_atan_small(double u) {
  const int N = 22;
  double u2 = u * u;
  double s = 1.0 / (double)(2 * N + 1);
#pragma unroll
  for (int n = N - 1; n >= 0; --n) s = 1.0 / (double)(2 * n + 1) - u2 * s;
  return u * s;
}
__device__ __forceinline__ double spec_atan01(double t) {
  const double kPi = 3.141592653589793;
  if (t > 0.41421356237309503) return kPi / 4.0 + spec_atan_small((t - 1.0) / (t + 1.0));
  return spec_atan_small(t);
}
__device__ __forceinline__ double spec_atan2_ypos(double y, double x) {
  const double kPi = 3.141592653589793;
  double ax = fabs(x);
  if (y == 0.0) return x < 0.0 ? kPi : 0.0;
  double r;
  if (ax >= y) r = spec_atan01(y / ax);
  else r = kPi / 2.0 - spec_atan01(ax / y);
  if (x < 0.0) r = kPi - r;
  return r;
}
__device__ __forceinline__ double spec_cos(double x) {
  double x2 = x * x, c = 1.0;
#pragma unroll
  for (int n = 13; n >= 1; --n) c = 1.0 - (x2 * (1.0 / (double)((2 * n - 1) * (2 * n)))) * c;
  return c;
}
__device__ __forceinline__ double spec_sin(double x) {
  double x2 = x * x, s = 1.0;
#pragma unroll
  for (int n = 13; n >= 1; --n) s = 1.0 - (x2 * (1.0 / (double)((2 * n) * (2 * n + 1)))) * s;
  return x * s;
}

__device__ __forceinline__ void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float bb = b * b;
  float d = (float)((double)bb - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

__device__ __forceinline__ void swapf(float& a, float& b) {
  float t = a;
  a = b;
  b = t;
}

__device__ void compute_roots(const float m[3][3], float roots[3]) {
  float c0 = m[0][0] * m[1][1] * m[2][2];
  c0 = c0 + 2.0f * m[0][1] * m[0][2] * m[1][2];
  c0 = c0 - m[0][0] * m[1][2] * m[1][2];
  c0 = c0 - m[1][1] * m[0][2] * m[0][2];
  c0 = c0 - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1];
  c1 = c1 - m[0][1] * m[0][1];
  c1 = c1 + m[0][0] * m[2][2];
  c1 = c1 - m[0][2] * m[0][2];
  c1 = c1 + m[1][1] * m[2][2];
  c1 = c1 - m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1];
  c2 = c2 + m[2][2];
  if (fabsf(c0) < 1.1920928955078125e-07f) {  // FLT_EPSILON
    compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = 1.7320508075688772f;  // sqrtf(3.0f)
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  float rho = sqrtf(-a_over_3);
  float sq = sqrtf(-q);
  float at = (float)spec_atan2_ypos((double)sq, (double)half_b);
  float theta = at * s_inv3;
  float cos_theta = (float)spec_cos((double)theta);
  float sin_theta = (float)spec_sin((double)theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    swapf(roots[1], roots[2]);
    if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
  }
  if (roots[0] <= 0.f) compute_roots2(c2, c1, roots);
}

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float sqnorm3(const float v[3]) {
  return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
}

// everything after the k-term sums: means, covariance, smallest eigenpair, flip, curvature
__device__ __forceinline__ float4 finish_normal(float a0, float a1, float a2, float a3, float a4, float a5, float a6,
                                                float a7, float a8, int cnt, const float4 self) {
  if (cnt < 3) {
    float nanv = __int_as_float(0x7FC00000);
    return make_float4(nanv, nanv, nanv, nanv);
  }
  float fc = (float)cnt;
  a0 /= fc; a1 /= fc; a2 /= fc; a3 /= fc; a4 /= fc; a5 /= fc; a6 /= fc; a7 /= fc; a8 /= fc;
  float cov[3][3];
  cov[0][0] = a0 - a6 * a6;
  cov[0][1] = a1 - a6 * a7;
  cov[0][2] = a2 - a6 * a8;
  cov[1][1] = a3 - a7 * a7;
  cov[1][2] = a4 - a7 * a8;
  cov[2][2] = a5 - a8 * a8;
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  float scale = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) scale = fmaxf(scale, fabsf(cov[r][cc]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;  // FLT_MIN
  float sm[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) sm[r][cc] = cov[r][cc] / scale;
  float roots[3];
  compute_roots(sm, roots);
  float eigenvalue = roots[0] * scale;
  sm[0][0] -= roots[0];
  sm[1][1] -= roots[0];
  sm[2][2] -= roots[0];
  float v1[3], v2[3], v3[3];
  cross3(sm[0], sm[1], v1);
  cross3(sm[0], sm[2], v2);
  cross3(sm[1], sm[2], v3);
  float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
  float nx, ny, nz;
  if (l1 >= l2 && l1 >= l3) {
    float s = sqrtf(l1);
    nx = v1[0] / s; ny = v1[1] / s; nz = v1[2] / s;
  } else if (l2 >= l1 && l2 >= l3) {
    float s = sqrtf(l2);
    nx = v2[0] / s; ny = v2[1] / s; nz = v2[2] / s;
  } else {
    float s = sqrtf(l3);
    nx = v3[0] / s; ny = v3[1] / s; nz = v3[2] / s;
  }
  float eig_sum = cov[0][0] + cov[1][1];
  eig_sum = eig_sum + cov[2][2];
  float curvature = (eig_sum != 0.f) ? fabsf(eigenvalue / eig_sum) : 0.f;
  float vx = 0.f - self.x, vy = 0.f - self.y, vz = 0.f - self.z;
  float ct = vx * nx + vy * ny;
  ct = ct + vz * nz;
  if (ct < 0.f) {
    nx *= -1.f; ny *= -1.f; nz *= -1.f;
  }
  return make_float4(nx, ny, nz, curvature);
}

// One warp per tile of 32 consecutive (Morton-adjacent) points, NRM_C neighbours at a time.
//   gather:     lane <-> (point, neighbour) ENTRY of the tile.  The 32 entries of a request are
//               the consecutive row entries of two points, so the index loads are two 64-byte
//               runs and the coordinate gathers fall on the few lines the neighbourhood of ONE
//               place occupies in Morton order; the coordinates go to shared memory.
//   accumulate: lane <-> POINT.  The nine FP32 sums advance in list order, exactly as
//               pcl::computeMeanAndCovarianceMatrix does — the order is part of the contract.
// The one-thread-per-point form of this kernel (rows read at a 4k-byte stride, every gather on
// its own line) ran at 97 % L1 tag throughput: 64 requests x 32 lines per warp
// (profiles/r01_misc_kernels_ncu.txt).
// Tile shape: C neighbours at a time, B gather rounds in flight per lane, W warps per block.
// (C, B, W) = (16, 8, 4) was round 1's; the variants are selectable with IFTWT_NRM_VARIANT for the
// tuning sweep (tools/tune.py), the default is the fastest measured on cfg3.
template <int C, int B, int W, int MINB = 0>
__global__ void __launch_bounds__(W * 32, MINB)
normals_kernel(const float4* __restrict__ spts, const u32* __restrict__ nbr, u32 n, int k,
               float4* __restrict__ out) {
  constexpr int STRIDE = C * 3 + 1;
  __shared__ float s_nb[W][32 * STRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 u0 = (blockIdx.x * W + warp) * 32u;
  if (u0 >= n) return;
  float* nb = s_nb[warp];
  const u32 i = u0 + lane;
  int cnt = 0;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
  const float nanv = __int_as_float(0x7FC00000);
  for (int c0 = 0; c0 < k; c0 += C) {
    const int cc = min(C, k - c0);
    for (int e0 = lane; e0 < 32 * C; e0 += 32 * B) {
      u32 v[B];
      float4 P[B];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int e = e0 + 32 * b, p = e / C, j = e % C;
        v[b] = (j < cc && u0 + p < n) ? nbr[(size_t)(u0 + p) * k + c0 + j] : SID_NONE;
      }
#pragma unroll
      for (int b = 0; b < B; ++b) P[b] = v[b] != SID_NONE ? spts[v[b]] : make_float4(nanv, 0.f, 0.f, 0.f);
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int e = e0 + 32 * b, p = e / C, j = e % C;
        float* d = nb + p * STRIDE + 3 * j;
        d[0] = P[b].x;
        d[1] = P[b].y;
        d[2] = P[b].z;
      }
    }
    __syncwarp();
    const float* mine = nb + lane * STRIDE;
#pragma unroll 4
    for (int j = 0; j < C; ++j) {
      const float px = mine[3 * j], py = mine[3 * j + 1], pz = mine[3 * j + 2];
      if (px != px) continue;  // missing neighbour (or past the row's end)
      a0 += px * px;
      a1 += px * py;
      a2 += px * pz;
      a3 += py * py;
      a4 += py * pz;
      a5 += pz * pz;
      a6 += px;
      a7 += py;
      a8 += pz;
      ++cnt;
    }
    __syncwarp();
  }
  if (i < n) out[i] = finish_normal(a0, a1, a2, a3, a4, a5, a6, a7, a8, cnt, spts[i]);
}

}  // namespace

int normals_run(iftwt_ctx* c) {
  if (c->knn_k == 0) return ctx_fail(c, IFTWT_ERR_STATE, "k-NN lists are not computed");
  if (c->nrm_k == c->knn_k) return IFTWT_OK;
  StageTimer st(c, IFTWT_STAGE_NORMALS);
  const u32 n = (u32)c->n;
  ENSURE(c, c->nrm, (size_t)n * 16);
  const int variant = getenv("IFTWT_NRM_VARIANT") ? atoi(getenv("IFTWT_NRM_VARIANT")) : 0;
#define NRM_LAUNCH(C, B, W, ...)                                                                                  \
  LAUNCH(c, (normals_kernel<C, B, W, ##__VA_ARGS__>), div_up(n, W * 32), W * 32, 0, c->spts.as<float4>(),        \
         c->nbr.as<u32>(), n, c->knn_k, c->nrm.as<float4>())
  // cfg3 on B200 (tools/tune.py, gpurun_out/r2_tune1.jsonl): (16,8,4) 1.146 ms, (16,4,4) 1.125, (8,8,4) 1.100,
  // (8,4,8) 1.208, (8,8,8) 1.078, (8,4,4) 1.218, (16,8,2) 1.400
  switch (variant) {
    case 1: NRM_LAUNCH(16, 4, 4); break;
    case 2: NRM_LAUNCH(8, 8, 4); break;
    case 3: NRM_LAUNCH(8, 4, 8); break;
    case 4: NRM_LAUNCH(16, 8, 4); break;
    case 5: NRM_LAUNCH(8, 4, 4); break;
    case 6: NRM_LAUNCH(16, 8, 2); break;
    case 7: NRM_LAUNCH(8, 8, 8, 6); break;
    case 8: NRM_LAUNCH(8, 4, 8, 8); break;
    case 9: NRM_LAUNCH(8, 8, 4, 12); break;
    case 10: NRM_LAUNCH(8, 8, 8, 1); break;
    case 11: NRM_LAUNCH(8, 8, 8, 5); break;
    default: NRM_LAUNCH(8, 8, 8, 5); break;  // 48 registers without the 4-byte spill of the unbounded build: 1.07 -> 0.99 ms
  }
#undef NRM_LAUNCH
  CHECK_LAUNCH(c);
  c->nrm_k = c->knn_k;
  c->seeds_valid = false;
  return IFTWT_OK;
}
