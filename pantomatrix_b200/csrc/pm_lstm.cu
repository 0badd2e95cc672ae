// Persistent bidirectional LSTM layer (CaMN / DisCo decoders, BASELINE configs[2],[3]).
// Replaces the cuDNN/ATen recurrence inside nn.LSTM (reference models/camn_audio/modeling_camn_audio.py:205-217,
// 264-271; models/disco_audio/modeling_disco_audio.py:212-216,255).  Contract: include/pm_emage.h.
//
// The input projections W_ih x + b for all time steps are one big tap-GEMM; this kernel runs the sequential part.
// One cooperative launch per layer: 2 * G CTAs, CTA (dir, slot) owns H/G hidden units of one direction and keeps
// their 4*H/G rows of W_hh (fp32) in shared memory for all T steps.  Per step every CTA reloads h_{t-1} of its
// direction (written by all G CTAs one step earlier, read with ld.global.cg), forms its gate pre-activations for
// the whole batch, updates its private cell states (registers) and publishes h_t; the G CTAs of a direction then
// meet at a global-memory barrier (release/acquire on a counter).  Math is exact-order fp32.
#include "pm_common.cuh"
#include "../../include/pm_emage.h"

namespace {

constexpr int LB = 64;        // batch rows per launch (threads = 4 * LB)
constexpr int NT = 256;

struct LstmParams {
  const float* xproj; long long x_bs; int ldx;   // (B, T, >= 2*4H): column dir*4H + gate*H + unit
  const float* whh;                              // (2, 4H, H)
  float* y; long long y_bs; int ldy;             // (B, T, >= 2H): forward h in [0,H), backward in [H,2H)
  unsigned int* barrier;                         // 2 counters, zero at launch
  int B, T, H, G;                                // G CTAs per direction, H % (G * 2) == 0... units per CTA = H / G (== 8)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(NT, 1) lstm_bidir_kernel(LstmParams p) {
  extern __shared__ float smem[];
  const int H = p.H, HP = H + 4;                // row pad: 16B aligned, conflict-free float4 rows
  const int UPC = H / p.G;                      // hidden units owned by this CTA (8)
  float* Ws = smem;                             // [4 * UPC][HP]  rows ordered gate-major: gate * UPC + unit_local
  float* hs = Ws + 4 * UPC * HP;                // [LB][HP]
  const int dir = blockIdx.x / p.G, slot = blockIdx.x % p.G;
  const int tid = threadIdx.x;
  const int b = tid >> 2, ug = tid & 3;         // batch row, unit pair (units 2*ug, 2*ug+1 of this CTA)
  const int u0 = slot * UPC;                    // first global unit of this CTA

  // resident W_hh slice
  const float* W = p.whh + (long long)dir * 4 * H * H;
  for (int i = tid; i < 4 * UPC * (H / 4); i += NT) {
    const int r = i / (H / 4), k4 = i % (H / 4);
    const int gate = r / UPC, ul = r % UPC;
    const float4 v = *reinterpret_cast<const float4*>(W + (long long)(gate * H + u0 + ul) * H + k4 * 4);
    *reinterpret_cast<float4*>(Ws + r * HP + k4 * 4) = v;
  }
  float c0 = 0.f, c1 = 0.f;                      // cell states of this thread's two units (batch row b)
  const bool active = b < p.B;
  const float* xrow = p.xproj + (long long)b * p.x_bs + (long long)dir * 4 * H + u0 + 2 * ug;
  float* yrow = p.y + (long long)b * p.y_bs + (long long)dir * H;
  __syncthreads();

  for (int s = 0; s < p.T; ++s) {
    const int t = dir == 0 ? s : p.T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;    // time index holding h_{prev}
    // gate pre-activations start from the input projection (issued early: global latency overlaps the h load)
    float acc[8];
    if (active) {
      const float* x = xrow + (long long)t * p.ldx;
#pragma unroll
      for (int g = 0; g < 4; ++g) { acc[2 * g] = __ldg(x + g * H); acc[2 * g + 1] = __ldg(x + g * H + 1); }
    }
    // h_{prev} of this direction for the whole batch -> smem (zeros at the first step)
    for (int i = tid; i < LB * (H / 4); i += NT) {
      const int r = i / (H / 4), k4 = i % (H / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s > 0 && r < p.B)
        v = __ldcg(reinterpret_cast<const float4*>(p.y + (long long)r * p.y_bs + (long long)tp * p.ldy + dir * H + k4 * 4));
      *reinterpret_cast<float4*>(hs + r * HP + k4 * 4) = v;
    }
    __syncthreads();
    if (active && s > 0) {
      const float* hb = hs + b * HP;
      const float* w = Ws + (2 * ug) * HP;       // rows: gate * UPC + 2*ug (+1)
#pragma unroll 2
      for (int k = 0; k < H; k += 4) {
        const float4 h4 = *reinterpret_cast<const float4*>(hb + k);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 wa = *reinterpret_cast<const float4*>(w + (g * UPC) * HP + k);
          const float4 wb = *reinterpret_cast<const float4*>(w + (g * UPC + 1) * HP + k);
          acc[2 * g] = fmaf(h4.x, wa.x, acc[2 * g]); acc[2 * g] = fmaf(h4.y, wa.y, acc[2 * g]);
          acc[2 * g] = fmaf(h4.z, wa.z, acc[2 * g]); acc[2 * g] = fmaf(h4.w, wa.w, acc[2 * g]);
          acc[2 * g + 1] = fmaf(h4.x, wb.x, acc[2 * g + 1]); acc[2 * g + 1] = fmaf(h4.y, wb.y, acc[2 * g + 1]);
          acc[2 * g + 1] = fmaf(h4.z, wb.z, acc[2 * g + 1]); acc[2 * g + 1] = fmaf(h4.w, wb.w, acc[2 * g + 1]);
        }
      }
    }
    if (active) {                                 // gates i, f, g, o -> c, h   (nn.LSTM equations)
      c0 = sigmoidf_(acc[2]) * c0 + sigmoidf_(acc[0]) * tanhf(acc[4]);
      c1 = sigmoidf_(acc[3]) * c1 + sigmoidf_(acc[1]) * tanhf(acc[5]);
      float* yo = yrow + (long long)t * p.ldy + u0 + 2 * ug;
      __stcg(yo, sigmoidf_(acc[6]) * tanhf(c0));
      __stcg(yo + 1, sigmoidf_(acc[7]) * tanhf(c1));
    }
    // direction-wide barrier: every CTA of this direction has published h_t
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicAdd(p.barrier + dir, 1u);
      const unsigned int target = (unsigned int)(s + 1) * (unsigned int)p.G;
      const long long t0 = clock64();
      while (*reinterpret_cast<volatile unsigned int*>(p.barrier + dir) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();          // never hang the GPU on a protocol bug
      }
      __threadfence();
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int pm_lstm_bidir_f32(const float* xproj, long long x_bs, int ldx, const float* whh,
                                 float* y, long long y_bs, int ldy, unsigned int* barrier,
                                 int batch, int t, int hidden, void* stream) {
  PM_REQUIRE(xproj && whh && y && barrier && batch > 0 && t > 0);
  if (hidden != 512) return PM_EUNSUPPORTED;
  PM_REQUIRE(ldx >= 8 * hidden && ldy >= 2 * hidden && (ldy & 3) == 0 && (y_bs & 3) == 0);
  PM_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(whh) & 15) == 0);
  const int G = 64;                                    // 128 CTAs: 8 units (32 gate rows, 66 KB of W_hh) each
  const size_t smem = (size_t)(4 * (hidden / G) + LB) * (hidden + 4) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lstm_bidir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < batch; b0 += LB) {             // batch rows are independent: chunks of 64
    cudaError_t e = cudaMemsetAsync(barrier, 0, 2 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    LstmParams p{xproj + (long long)b0 * x_bs, x_bs, ldx, whh, y + (long long)b0 * y_bs, y_bs, ldy, barrier,
                 batch - b0 < LB ? batch - b0 : LB, t, hidden, G};
    void* args[] = {&p};
    e = cudaLaunchCooperativeKernel((const void*)lstm_bidir_kernel, dim3(2 * G), dim3(NT), args, smem, st);
    if (e != cudaSuccess) return (int)e;
  }
  return PM_OK;
}
