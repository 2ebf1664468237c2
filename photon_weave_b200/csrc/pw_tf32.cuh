// Error-compensated TF32 (3 x TF32) complex products on the warp-level tensor path
// (mma.sync.m16n8k8), shared by the two-sided kernels (pw_blocks.cu) and the dense one-sided
// kernel (pw_dense.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pw {

// ---- complex64 on the TF32 tensor cores, error-compensated (3 x TF32) ---------------------------
// The FP64 tensor-core path gives complex64 the complex128 TIMES (it is DMMA-bound: 37 TFLOP/s
// measured), i.e. 35-46 % of the complex64 HBM roof.  mma.sync.m16n8k8 TF32 runs at 279 TFLOP/s
// (measured); splitting every operand into hi + lo TF32 parts and accumulating
// a_lo b_hi + a_hi b_lo + a_hi b_hi in FP32 keeps ~2^-21 relative accuracy (complex64 tolerance:
// 1e-5) at a third of that rate -- still 2.5x the FP64 tensor cores.  The fragment conventions
// coincide with the complex FP64 product: with A = [Wr; Wi] (16 x 8) the registers a0..a3 of
// lane (gid, tig) are exactly (w0.x, w0.y, w1.x, w1.y), B = Xr is (x0.x, x1.x), and the
// accumulators c0..c3 are (Yr[gid][2 tig], Yr[gid][2 tig + 1], Yi[..], Yi[..]): a second MMA with
// A = [-Wi; Wr], B = Xi completes the complex product -- 2 MMAs instead of 8 DMMAs.
// hi = x with the 13 low mantissa bits cleared (what the TF32 datapath reads of a raw FP32 word
// anyway), lo = x - hi (exact in FP32; the hardware truncates it again on use: 2^-20 relative
// overall).  One AND and one FADD per value -- cvt.rna.tf32 is a conversion-pipe instruction
// and made the split as expensive as the products it feeds (measured: no gain over FP64).
struct Tf2 { uint32_t hi, lo; };
__device__ __forceinline__ Tf2 tf_split(const float x) {
  Tf2 r;
  r.hi = __float_as_uint(x) & 0xffffe000u;
  r.lo = __float_as_uint(x - __uint_as_float(r.hi));
  return r;
}
__device__ __forceinline__ void mma_tf32(float& c0, float& c1, float& c2, float& c3, const uint32_t a0,
                                         const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                         const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c0), "+f"(c1), "+f"(c2), "+f"(c3) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// operand fragment of a complex 8 x 8 matrix, pre-split: [re; im] rows and [-im; re] rows
struct TfW {
  Tf2 re0, im0, re1, im1, nim0, nim1;   // w0.x, w0.y, w1.x, w1.y, -w0.y, -w1.y
};
__device__ __forceinline__ TfW tf_w(const float2 w0, const float2 w1) {
  TfW r;
  r.re0 = tf_split(w0.x); r.im0 = tf_split(w0.y); r.re1 = tf_split(w1.x); r.im1 = tf_split(w1.y);
  r.nim0 = tf_split(-w0.y); r.nim1 = tf_split(-w1.y);
  return r;
}
// Y = W X (complex 8 x 8 times 8 x 8): lane (gid, tig) holds W[gid][tig], W[gid][4 + tig] (in w),
// x0 = X[tig][n = gid], x1 = X[4 + tig][gid]; y0 = Y[gid][2 tig], y1 = Y[gid][2 tig + 1]
__device__ __forceinline__ void cprod_tf32(const TfW& w, const float2 x0, const float2 x1, float2& y0, float2& y1) {
  const Tf2 xr0 = tf_split(x0.x), xr1 = tf_split(x1.x), xi0 = tf_split(x0.y), xi1 = tf_split(x1.y);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
  // small terms first
  mma_tf32(c0, c1, c2, c3, w.re0.lo, w.im0.lo, w.re1.lo, w.im1.lo, xr0.hi, xr1.hi);
  mma_tf32(c0, c1, c2, c3, w.nim0.lo, w.re0.lo, w.nim1.lo, w.re1.lo, xi0.hi, xi1.hi);
  mma_tf32(c0, c1, c2, c3, w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0.lo, xr1.lo);
  mma_tf32(c0, c1, c2, c3, w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0.lo, xi1.lo);
  mma_tf32(c0, c1, c2, c3, w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0.hi, xr1.hi);
  mma_tf32(c0, c1, c2, c3, w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0.hi, xi1.hi);
  y0 = make_float2(c0, c2);
  y1 = make_float2(c1, c3);
}
// the TRANSPOSED product Y^T = X^T W^T: A = [Xr^T; Xi^T] (fibers x inputs, real rows then
// imaginary rows), B = Wr^T, then B = Wi^T.  Lane (gid, tig): x0 = X^T[fiber gid][tig],
// x1 = X^T[gid][4 + tig]; w0 = W[gid][tig] = W^T[tig][gid], w1 = W[gid][4 + tig];
// y0 = Y^T[fiber gid][2 tig], y1 = Y^T[gid][2 tig + 1].
__device__ __forceinline__ void cprod_tf32_t(const TfW& w, const float2 x0, const float2 x1, float2& y0, float2& y1) {
  const Tf2 xr0 = tf_split(x0.x), xr1 = tf_split(x1.x), xi0 = tf_split(x0.y), xi1 = tf_split(x1.y);
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;   // [Xr Wr^T; Xi Wr^T]
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;   // [Xr Wi^T; Xi Wi^T]
  mma_tf32(p0, p1, p2, p3, xr0.lo, xi0.lo, xr1.lo, xi1.lo, w.re0.hi, w.re1.hi);
  mma_tf32(q0, q1, q2, q3, xr0.lo, xi0.lo, xr1.lo, xi1.lo, w.im0.hi, w.im1.hi);
  mma_tf32(p0, p1, p2, p3, xr0.hi, xi0.hi, xr1.hi, xi1.hi, w.re0.lo, w.re1.lo);
  mma_tf32(q0, q1, q2, q3, xr0.hi, xi0.hi, xr1.hi, xi1.hi, w.im0.lo, w.im1.lo);
  mma_tf32(p0, p1, p2, p3, xr0.hi, xi0.hi, xr1.hi, xi1.hi, w.re0.hi, w.re1.hi);
  mma_tf32(q0, q1, q2, q3, xr0.hi, xi0.hi, xr1.hi, xi1.hi, w.im0.hi, w.im1.hi);
  y0 = make_float2(p0 - q2, q0 + p2);   // Yr = Xr Wr^T - Xi Wi^T, Yi = Xr Wi^T + Xi Wr^T
  y1 = make_float2(p1 - q3, q1 + p3);
}

// accumulating form: c0..c3 += (Yr[gid][2 tig], Yr[gid][2 tig + 1], Yi[gid][2 tig], Yi[gid][2 tig + 1])
__device__ __forceinline__ void cprod_tf32_acc(const TfW& w, const float2 x0, const float2 x1, float (&c)[4]) {
  const Tf2 xr0 = tf_split(x0.x), xr1 = tf_split(x1.x), xi0 = tf_split(x0.y), xi1 = tf_split(x1.y);
  mma_tf32(c[0], c[1], c[2], c[3], w.re0.lo, w.im0.lo, w.re1.lo, w.im1.lo, xr0.hi, xr1.hi);
  mma_tf32(c[0], c[1], c[2], c[3], w.nim0.lo, w.re0.lo, w.nim1.lo, w.re1.lo, xi0.hi, xi1.hi);
  mma_tf32(c[0], c[1], c[2], c[3], w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0.lo, xr1.lo);
  mma_tf32(c[0], c[1], c[2], c[3], w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0.lo, xi1.lo);
  mma_tf32(c[0], c[1], c[2], c[3], w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0.hi, xr1.hi);
  mma_tf32(c[0], c[1], c[2], c[3], w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0.hi, xi1.hi);
}

// NB independent items that share the operand fragment w, issued TERM-major: consecutive MMAs in
// program order belong to different accumulators (a warp issues in order; six dependent MMAs per
// item back to back would serialise on the tensor-pipe latency).  ACC: add to c instead of
// starting from zero.
template <int NB, bool ACC>
__device__ __forceinline__ void cprod_tf32_batch(const TfW& w, const float2* x0, const float2* x1, float (*c)[4]) {
  Tf2 xr0[NB], xr1[NB], xi0[NB], xi1[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    xr0[b] = tf_split(x0[b].x); xr1[b] = tf_split(x1[b].x); xi0[b] = tf_split(x0[b].y); xi1[b] = tf_split(x1[b].y);
    if (!ACC) { c[b][0] = 0.f; c[b][1] = 0.f; c[b][2] = 0.f; c[b][3] = 0.f; }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.re0.lo, w.im0.lo, w.re1.lo, w.im1.lo, xr0[b].hi, xr1[b].hi);
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.nim0.lo, w.re0.lo, w.nim1.lo, w.re1.lo, xi0[b].hi, xi1[b].hi);
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0[b].lo, xr1[b].lo);
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0[b].lo, xi1[b].lo);
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.re0.hi, w.im0.hi, w.re1.hi, w.im1.hi, xr0[b].hi, xr1[b].hi);
#pragma unroll
  for (int b = 0; b < NB; ++b) mma_tf32(c[b][0], c[b][1], c[b][2], c[b][3], w.nim0.hi, w.re0.hi, w.nim1.hi, w.re1.hi, xi0[b].hi, xi1[b].hi);
}

}  // namespace pw
