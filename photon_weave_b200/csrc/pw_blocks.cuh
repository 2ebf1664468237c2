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
  int inplace_safe;  // 1: no block writes a position a DIFFERENT block still has to read
  int seg_handled;   // usable && inplace_safe: the staged (segmented) kernel took it
  int pad[2];
};

// device pointers of one analysed operator
struct BlockArrays {
  BlockDescHeader* hdr = nullptr;
  int4* blocks = nullptr;        // per block: nr, nc, offset into offs, offset into vals
  int64_t* offs = nullptr;       // per block: nc column offsets then nr row offsets (elements)
  int64_t* zero_offs = nullptr;  // offsets of all-zero rows
  void* vals = nullptr;          // per block: dense nr x nc values (row-major)
};

struct BlockDesc {
  Scratch mem;
  BlockArrays arr;
  BlockDescHeader* hdr = nullptr;
  bool staged = false;  // launch_blocks_seg was used: the flag to honour is seg_handled
  const int* skip_flag() const {
    return hdr ? (staged ? &hdr->seg_handled : &hdr->usable) : nullptr;
  }
};

// Enqueues the structure analysis of a dense t x t operator (t <= kBlockMaxT); element
// offsets come from `tm` (global strides for the direct kernel, tile strides for the
// tiled kernel).
int build_block_desc(const void* op, uint32_t t, bool conj_op, const TargetMap& tm, int dtype,
                     Scratch* mem, BlockArrays* arr, cudaStream_t stream);

// Enqueues structure analysis + the block apply kernels.  `conj_op`: use conj(op).
// Requires in != out.  After this call the caller must enqueue a fallback path whose
// kernels exit when *desc->skip_flag() != 0.
int launch_blocks(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                  bool conj_op, BlockDesc* desc, cudaStream_t stream);

// Staged variant for layouts whose innermost axis is a target: a warp stages 32 adjacent
// fibers in shared memory with coalesced cp.async rows, applies the blocks in place (one
// fiber per lane), and streams the rows back.  Runs when the analysis reports
// usable && inplace_safe (desc->hdr->seg_handled); the fallback must take that as skip flag.
// Returns PW_ERR_UNSUPPORTED when the geometry does not fit.
int launch_blocks_seg(const void* in, void* out, const void* op, const int64_t* axes, int naxes,
                      const int* targets, int nt, int dtype, bool conj_op, BlockDesc* desc,
                      cudaStream_t stream);

}  // namespace pw
