// Block-structured operator descriptor (see pw_blocks.cu).
#pragma once
#include "pw_internal.cuh"

namespace pw {

constexpr int kBlockMaxN = 8;     // largest dense block the register kernel takes
constexpr uint32_t kBlockMaxT = 128;  // largest operator dimension analysed on the device

struct BlockDescHeader {
  int nblocks;
  int maxn;    // largest block side
  int nzero;   // all-zero rows (outputs that are simply cleared)
  int usable;  // 1: the block kernel handles this operator; fallback kernels must exit
};

struct BlockDesc {
  Scratch mem;
  BlockDescHeader* hdr = nullptr;
  const int* skip_flag() const { return hdr ? &hdr->usable : nullptr; }
};

// Enqueues structure analysis + the block apply kernels.  `conj_op`: use conj(op).
// Requires in != out.  After this call the caller must enqueue a fallback path whose
// kernels exit when *desc->skip_flag() != 0.
int launch_blocks(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                  bool conj_op, BlockDesc* desc, cudaStream_t stream);

}  // namespace pw
