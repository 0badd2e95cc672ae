// Whole-sequence multi-head attention for the EMAGE transformer layers (fp32 SIMT engine).
// T <= 64 tokens, head_dim = 192, no masks: the full score tile lives on chip, so there is no
// online-softmax pass.  One CTA per (clip, head).  Contract: include/pm_emage.h (pm_attention_f32).
#include <stdlib.h>
#include "pm_common.cuh"
#include "../../include/pm_emage.h"

namespace {

constexpr int TMAX = 64;
constexpr int HD = 192;
constexpr int HDP = HD + 1;     // +1 float: conflict-free column walks over rows
constexpr int NT = 256;

template <bool F16>
__global__ void __launch_bounds__(NT) attention_f32_kernel(
    const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
    const float* __restrict__ V, int ldv, float* __restrict__ O, int ldo,
    int heads, int tq, int tk, float scale, PmPlanes P) {
  extern __shared__ float smem[];
  float* Qs = smem;                    // [TMAX][HDP]
  float* Ks = Qs + TMAX * HDP;         // [TMAX][HDP]
  float* Vs = Ks + TMAX * HDP;         // [TMAX][HD]
  float* S = Vs + TMAX * HD;           // [TMAX][TMAX+1]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x;

  // stage Q, K, V head slices (float4 global loads, scalar smem stores because of the +1 padding)
  for (int i = tid; i < TMAX * (HD / 4); i += NT) {
    const int r = i / (HD / 4), c4 = i % (HD / 4);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
    if (r < tq) q = *reinterpret_cast<const float4*>(Q + (long long)(b * tq + r) * ldq + h * HD + c4 * 4);
    if (r < tk) {
      k = *reinterpret_cast<const float4*>(K + (long long)(b * tk + r) * ldk + h * HD + c4 * 4);
      v = *reinterpret_cast<const float4*>(V + (long long)(b * tk + r) * ldv + h * HD + c4 * 4);
    }
    float* qd = Qs + r * HDP + c4 * 4;
    qd[0] = q.x; qd[1] = q.y; qd[2] = q.z; qd[3] = q.w;
    float* kd = Ks + r * HDP + c4 * 4;
    kd[0] = k.x; kd[1] = k.y; kd[2] = k.z; kd[3] = k.w;
    *reinterpret_cast<float4*>(Vs + r * HD + c4 * 4) = v;
  }
  __syncthreads();

  // S = scale * Q K^T : each thread a 4x4 block of the 64x64 tile
  {
    const int ti = tid >> 4, tj = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const float* q0 = Qs + (ti * 4) * HDP;
    const float* k0 = Ks + (tj * 4) * HDP;
#pragma unroll 4
    for (int d = 0; d < HD; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { qv[a] = q0[a * HDP + d]; kv[a] = k0[a * HDP + d]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(qv[a], kv[c], acc[a][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) S[(ti * 4 + a) * (TMAX + 1) + tj * 4 + c] = acc[a][c] * scale;
  }
  __syncthreads();

  // row softmax over the tk valid keys: warp w owns rows w, w+8, ...
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < tq; r += NT / 32) {
      float* row = S + r * (TMAX + 1);
      const float v0 = lane < tk ? row[lane] : -INFINITY;
      const float v1 = lane + 32 < tk ? row[lane + 32] : -INFINITY;
      const float m = pm_warp_max(fmaxf(v0, v1));
      const float e0 = lane < tk ? expf(v0 - m) : 0.f;
      const float e1 = lane + 32 < tk ? expf(v1 - m) : 0.f;
      const float inv = 1.f / pm_warp_sum(e0 + e1);
      row[lane] = e0 * inv;
      row[lane + 32] = e1 * inv;
    }
  }
  __syncthreads();

  // O = P V : thread -> 4 rows x 12 strided columns
  {
    const int tr = tid >> 4, tc = tid & 15;
    float acc[4][12];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int m = 0; m < 12; ++m) acc[a][m] = 0.f;
    for (int j = 0; j < tk; ++j) {
      float pv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) pv[a] = S[(tr * 4 + a) * (TMAX + 1) + j];
#pragma unroll
      for (int m = 0; m < 12; ++m) {
        const float vv = Vs[j * HD + tc + 16 * m];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][m] = fmaf(pv[a], vv, acc[a][m]);
      }
    }
    // stage the 64 x 192 output tile in smem (the Q region is dead by now) so global writes are row-contiguous
    __syncthreads();
    float* Os = Qs;                      // [TMAX][HD]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int m = 0; m < 12; ++m) Os[(tr * 4 + a) * HD + tc + 16 * m] = acc[a][m];
  }
  __syncthreads();
  {
    const float* Os = Qs;
    const bool vec_p = P.ptr && ((P.ld & 3) == 0) && ((P.ps & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.ptr) & 7) == 0);
    for (int i = tid; i < tq * (HD / 4); i += NT) {
      const int r = i / (HD / 4), c4 = i % (HD / 4);
      const float4 v = *reinterpret_cast<const float4*>(Os + r * HD + c4 * 4);
      const long long row = (long long)b * tq + r;
      if (O) *reinterpret_cast<float4*>(O + row * ldo + h * HD + c4 * 4) = v;
      if (P.ptr) {
        if (vec_p) pm_store_planes4_t<F16>(P, row, h * HD + c4 * 4, v);
        else {
          pm_store_planes_t<F16>(P, row, h * HD + c4 * 4, v.x); pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 1, v.y);
          pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 2, v.z); pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 3, v.w);
        }
      }
    }
  }
}

constexpr size_t kSmemBytes = (size_t)(2 * TMAX * HDP + TMAX * HD + TMAX * (TMAX + 1)) * sizeof(float);

// ---------------------------------------------------------------------------------------------------
// Tensor-core variant (opt-in, PM_ATTN_MMA=1; not yet measured): same contract, QK^T and PV on mma.sync
// m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, products lo*hi + hi*lo + hi*hi, fp32 accumulate), which keeps
// fp32-class accuracy (~2^-21).  The fp32 SIMT kernel above is shared-memory-bandwidth bound (8 scalar LDS per
// 16 FMA in the score phase); here a warp loads each A fragment once per k-step and reuses it over 4-12 column tiles.
// The op is tiny (64 x 64 x 192 per head) and latency bound, so the legacy warp-level MMA is the pragmatic tool;
// the GEMMs that carry the FLOPs are on tcgen05.
constexpr int QP = HD + 4;      // Q / K row stride (floats): 16-byte rows, conflict-free fragment loads (196 % 32 == 4)
constexpr int VP = HD + 8;      // V row stride: (k * 200 + n) % 32 distinct over a B fragment (200 % 32 == 8)
constexpr int PP = TMAX + 4;    // score / probability row stride

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <bool F16>
__global__ void __launch_bounds__(NT) attention_mma_kernel(
    const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk,
    const float* __restrict__ V, int ldv, float* __restrict__ O, int ldo,
    int heads, int tq, int tk, float scale, PmPlanes P) {
  extern __shared__ float smem[];
  float* Qs = smem;                    // [TMAX][QP]   (later: the output tile [TMAX][HD])
  float* Ks = Qs + TMAX * QP;          // [TMAX][QP]
  float* Vs = Ks + TMAX * QP;          // [TMAX][VP]
  float* Ps = Vs + TMAX * VP;          // [TMAX][PP]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;                 // fragment coordinates (PTX m16n8k8 layouts)

  for (int i = tid; i < TMAX * (HD / 4); i += NT) {
    const int r = i / (HD / 4), c4 = i % (HD / 4);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
    if (r < tq) q = *reinterpret_cast<const float4*>(Q + (long long)(b * tq + r) * ldq + h * HD + c4 * 4);
    if (r < tk) {
      k = *reinterpret_cast<const float4*>(K + (long long)(b * tk + r) * ldk + h * HD + c4 * 4);
      v = *reinterpret_cast<const float4*>(V + (long long)(b * tk + r) * ldv + h * HD + c4 * 4);
    }
    *reinterpret_cast<float4*>(Qs + r * QP + c4 * 4) = q;
    *reinterpret_cast<float4*>(Ks + r * QP + c4 * 4) = k;
    *reinterpret_cast<float4*>(Vs + r * VP + c4 * 4) = v;
  }
  __syncthreads();

  const int mt = warp & 3, nh = warp >> 2;               // 16-row tile, column half
  // ---- S = scale * Q K^T: warp tile 16 x 32 (4 column tiles of 8)
  {
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
    const float* qa = Qs + (16 * mt + g) * QP + t;
    const float* kb = Ks + (32 * nh + g) * QP + t;
#pragma unroll 2
    for (int k0 = 0; k0 < HD; k0 += 8) {
      uint32_t ah[4], al[4];
      split_tf32(qa[k0], ah[0], al[0]);
      split_tf32(qa[8 * QP + k0], ah[1], al[1]);
      split_tf32(qa[k0 + 4], ah[2], al[2]);
      split_tf32(qa[8 * QP + k0 + 4], ah[3], al[3]);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        uint32_t bh[2], bl[2];
        split_tf32(kb[(8 * n) * QP + k0], bh[0], bl[0]);
        split_tf32(kb[(8 * n) * QP + k0 + 4], bh[1], bl[1]);
        mma_tf32(acc[n], al, bh);
        mma_tf32(acc[n], ah, bl);
        mma_tf32(acc[n], ah, bh);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      float* s0 = Ps + (16 * mt + g) * PP + 32 * nh + 8 * n + 2 * t;
      s0[0] = acc[n][0] * scale; s0[1] = acc[n][1] * scale;
      s0[8 * PP] = acc[n][2] * scale; s0[8 * PP + 1] = acc[n][3] * scale;
    }
  }
  __syncthreads();

  // ---- row softmax over the tk valid keys (warp w owns rows w, w+8, ...); padded keys get probability 0
  for (int r = warp; r < TMAX; r += NT / 32) {
    float* row = Ps + r * PP;
    const float v0 = lane < tk ? row[lane] : -INFINITY;
    const float v1 = lane + 32 < tk ? row[lane + 32] : -INFINITY;
    const float m = pm_warp_max(fmaxf(v0, v1));
    const float e0 = lane < tk ? expf(v0 - m) : 0.f;
    const float e1 = lane + 32 < tk ? expf(v1 - m) : 0.f;
    const float inv = 1.f / pm_warp_sum(e0 + e1);
    row[lane] = e0 * inv;
    row[lane + 32] = e1 * inv;
  }
  __syncthreads();

  // ---- O = P V: warp tile 16 x 96 (12 column tiles of 8), K = 64 keys
  {
    float acc[12][4];
#pragma unroll
    for (int n = 0; n < 12; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[n][c] = 0.f;
    const float* pa = Ps + (16 * mt + g) * PP + t;
    const float* vb = Vs + t * VP + 96 * nh + g;
#pragma unroll 1
    for (int k0 = 0; k0 < TMAX; k0 += 8) {
      uint32_t ah[4], al[4];
      split_tf32(pa[k0], ah[0], al[0]);
      split_tf32(pa[8 * PP + k0], ah[1], al[1]);
      split_tf32(pa[k0 + 4], ah[2], al[2]);
      split_tf32(pa[8 * PP + k0 + 4], ah[3], al[3]);
#pragma unroll
      for (int n = 0; n < 12; ++n) {
        uint32_t bh[2], bl[2];
        split_tf32(vb[k0 * VP + 8 * n], bh[0], bl[0]);             // B[k][n] = V[key k0 + t][96 nh + 8 n + g]
        split_tf32(vb[(k0 + 4) * VP + 8 * n], bh[1], bl[1]);
        mma_tf32(acc[n], al, bh);
        mma_tf32(acc[n], ah, bl);
        mma_tf32(acc[n], ah, bh);
      }
    }
    __syncthreads();                     // everyone is done reading Q (score phase) - its region becomes the output tile
    float* Os = Qs;                      // [TMAX][HD]
#pragma unroll
    for (int n = 0; n < 12; ++n) {
      float* o0 = Os + (16 * mt + g) * HD + 96 * nh + 8 * n + 2 * t;
      *reinterpret_cast<float2*>(o0) = make_float2(acc[n][0], acc[n][1]);
      *reinterpret_cast<float2*>(o0 + 8 * HD) = make_float2(acc[n][2], acc[n][3]);
    }
  }
  __syncthreads();
  {
    const float* Os = Qs;
    const bool vec_p = P.ptr && ((P.ld & 3) == 0) && ((P.ps & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.ptr) & 7) == 0);
    for (int i = tid; i < tq * (HD / 4); i += NT) {
      const int r = i / (HD / 4), c4 = i % (HD / 4);
      const float4 v = *reinterpret_cast<const float4*>(Os + r * HD + c4 * 4);
      const long long row = (long long)b * tq + r;
      if (O) *reinterpret_cast<float4*>(O + row * ldo + h * HD + c4 * 4) = v;
      if (P.ptr) {
        if (vec_p) pm_store_planes4_t<F16>(P, row, h * HD + c4 * 4, v);
        else {
          pm_store_planes_t<F16>(P, row, h * HD + c4 * 4, v.x); pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 1, v.y);
          pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 2, v.z); pm_store_planes_t<F16>(P, row, h * HD + c4 * 4 + 3, v.w);
        }
      }
    }
  }
}

constexpr size_t kSmemBytesMma = (size_t)(2 * TMAX * QP + TMAX * VP + TMAX * PP) * sizeof(float);

}  // namespace

extern "C" int pm_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                                float* O, int ldo, int batch, int heads, int tq, int tk, int head_dim,
                                uint16_t* planes, long long p_ps, int p_ld, int p_nsplit, void* stream) {
  PM_REQUIRE(Q && K && V && (O || planes) && batch >= 0 && heads > 0);
  PM_TAKE_FMT(p_nsplit, f16);
  PM_REQUIRE(pm_planes_ok(planes, p_ps, p_ld, p_nsplit, heads * head_dim, false));
  const PmPlanes P{reinterpret_cast<__nv_bfloat16*>(planes), p_ps, p_ld, p_nsplit};
  if (head_dim != HD || tq > TMAX || tk > TMAX || tq <= 0 || tk <= 0) return PM_EUNSUPPORTED;
  PM_REQUIRE((ldq & 3) == 0 && (ldk & 3) == 0 && (ldv & 3) == 0 && (!O || (ldo & 3) == 0));
  if (batch == 0) return PM_OK;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  static const bool use_mma = getenv("PM_ATTN_MMA") && atoi(getenv("PM_ATTN_MMA")) != 0;
  if (use_mma) {
    static bool configured_mma = false;
    if (!configured_mma) {
      cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kSmemBytesMma);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesMma);
      if (e != cudaSuccess) return (int)e;
      configured_mma = true;
    }
    const float sc = 1.0f / sqrtf((float)head_dim);
    if (f16) attention_mma_kernel<true><<<batch * heads, NT, kSmemBytesMma, (cudaStream_t)stream>>>(
        Q, ldq, K, ldk, V, ldv, O, ldo, heads, tq, tk, sc, P);
    else attention_mma_kernel<false><<<batch * heads, NT, kSmemBytesMma, (cudaStream_t)stream>>>(
        Q, ldq, K, ldk, V, ldv, O, ldo, heads, tq, tk, sc, P);
    PM_LAUNCH_CHECK();
  }
  if (f16) attention_f32_kernel<true><<<batch * heads, NT, kSmemBytes, (cudaStream_t)stream>>>(
      Q, ldq, K, ldk, V, ldv, O, ldo, heads, tq, tk, 1.0f / sqrtf((float)head_dim), P);
  else attention_f32_kernel<false><<<batch * heads, NT, kSmemBytes, (cudaStream_t)stream>>>(
      Q, ldq, K, ldk, V, ldv, O, ldo, heads, tq, tk, 1.0f / sqrtf((float)head_dim), P);
  PM_LAUNCH_CHECK();
}
