// Arithmetic-peak probes: the SECOND roof of the dense-operator kernels (SURVEY.md section 8d;
// BASELINE.md section 2 left the FP64 peak "unmeasured").  Register-only kernels with long chains
// of independent FP64 FMAs / mma.sync.m8n8k4.f64 (DMMA) / FP32 FMAs; the caller times them with
// CUDA events (tools/fp64_peak.py) and the flop count is returned, so that dense 36x36 / 64x64
// operators can be quoted against a MEASURED compute roof as well as against HBM.
#include "pw_internal.cuh"

namespace pw {

constexpr int kProbeBlock = 256;
constexpr int kProbeCtas = 8;  // per SM

__global__ void __launch_bounds__(kProbeBlock)
k_probe_dfma(const int iters, double* __restrict__ sink) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1e-3 * (threadIdx.x + i);
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) sink[0] = s;  // keeps the chains alive, never true
}

__global__ void __launch_bounds__(kProbeBlock)
k_probe_sfma(const int iters, double* __restrict__ sink) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1e-3f * (threadIdx.x + i);
  const float m = 1.00001f, c = 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678f) sink[0] = s;
}

__global__ void __launch_bounds__(kProbeBlock)
k_probe_dmma(const int iters, double* __restrict__ sink) {
  double c0[8], c1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c0[i] = 0; c1[i] = 0; }
  const double a = 1e-3 * (threadIdx.x & 31), b = 1.0 + 1e-6 * (threadIdx.x >> 5);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) sink[0] = s;
}

// DMMA interleaved with scalar adds: 6 DMMAs + NADD independent adds per group (FP64 adds: the
// price of Gauss' three-product complex multiply on the FP64 tensor cores; FP32 adds: control)
template <int NADD, bool F64>
__global__ void __launch_bounds__(kProbeBlock)
k_probe_dmma_mix(const int iters, double* __restrict__ sink) {
  double c0[6], c1[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { c0[i] = 0; c1[i] = 0; }
  double ad[NADD];
  float af[NADD];
#pragma unroll
  for (int i = 0; i < NADD; ++i) { ad[i] = 1e-3 * (threadIdx.x + i); af[i] = (float)ad[i]; }
  const double a = 1e-3 * (threadIdx.x & 31), b = 1.0 + 1e-6 * (threadIdx.x >> 5);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
      if (i < NADD) {
        if (F64) asm volatile("add.f64 %0, %0, %1;" : "+d"(ad[i]) : "d"(b));
        else     asm volatile("add.f32 %0, %0, %1;" : "+f"(af[i]) : "f"(1.5f));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NADD; ++i) s += ad[i] + af[i];
  if (s == 12345.678) sink[0] = s;
}

// legacy warp-level TF32 tensor path (mma.sync m16n8k8): what an error-compensated 3 x TF32
// complex64 kernel would run on without tcgen05 / TMEM plumbing
__global__ void __launch_bounds__(kProbeBlock)
k_probe_tf32(const int iters, double* __restrict__ sink) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
  const uint32_t a0 = __float_as_uint(1e-3f * (threadIdx.x & 31)), a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
  const uint32_t b0 = __float_as_uint(1.0f + 1e-3f * (threadIdx.x >> 5)), b1 = b0 + 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678f) sink[0] = s;
}

}  // namespace pw

using namespace pw;

// kind 0: FP64 FMA, 1: FP64 mma.sync m8n8k4 (DMMA), 2: FP32 FMA, 3: TF32 mma.sync m16n8k8,
// 4..7: DMMA interleaved with scalar adds (MMA flops only are counted).  Enqueues ONE kernel of `iters`
// iterations on `stream`; *flops receives the real flops it executes.  `sink` is one device double.
extern "C" int pw_probe_arith(int kind, int iters, double* sink, double* flops, void* stream) {
  if (!sink || !flops || iters < 1 || kind < 0 || kind > 7) return PW_ERR_INVALID;
  const int grid = kNumSMs * kProbeCtas;
  cudaStream_t s = (cudaStream_t)stream;
  const double threads = (double)grid * kProbeBlock;
  if (kind == 0) {
    k_probe_dfma<<<grid, kProbeBlock, 0, s>>>(iters, sink);
    *flops = threads * 16.0 * 2.0 * iters;
  } else if (kind == 2) {
    k_probe_sfma<<<grid, kProbeBlock, 0, s>>>(iters, sink);
    *flops = threads * 16.0 * 2.0 * iters;
  } else if (kind == 3) {
    k_probe_tf32<<<grid, kProbeBlock, 0, s>>>(iters, sink);
    *flops = (threads / 32.0) * 8.0 * 2048.0 * iters;  // 8 MMAs of 2*16*8*8 flops per warp iteration
  } else if (kind >= 4) {
    // 4: 6 DMMA + 1 FP64 add, 5: + 2 FP64 adds, 6: + 3 FP64 adds, 7: + 2 FP32 adds; MMA flops only
    if (kind == 4) k_probe_dmma_mix<1, true><<<grid, kProbeBlock, 0, s>>>(iters, sink);
    else if (kind == 5) k_probe_dmma_mix<2, true><<<grid, kProbeBlock, 0, s>>>(iters, sink);
    else if (kind == 6) k_probe_dmma_mix<3, true><<<grid, kProbeBlock, 0, s>>>(iters, sink);
    else k_probe_dmma_mix<2, false><<<grid, kProbeBlock, 0, s>>>(iters, sink);
    *flops = (threads / 32.0) * 6.0 * 512.0 * iters;
  } else {
    k_probe_dmma<<<grid, kProbeBlock, 0, s>>>(iters, sink);
    *flops = (threads / 32.0) * 8.0 * 512.0 * iters;  // 8 MMAs of 2*8*8*4 flops per warp iteration
  }
  PW_LAUNCHED();
  return PW_OK;
}
