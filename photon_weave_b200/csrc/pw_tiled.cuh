// Tiled apply path: TMA-staged shared-memory tiles + CSR operator.
//
// The state tensor is described to the TMA unit as a rank<=5 tensor whose axes are
// the merged runs of target / non-target subsystems (never transposed in HBM).  A
// CTA loops over tiles (persistent, double-buffered cp.async.bulk.tensor loads with
// mbarrier completion), runs one or more PASSES over the tile in shared memory
// (each pass = one sparse operator acting along some of the tile's axes) and writes
// the result back with a TMA store.  Density matrices get O rho O^dagger / sum_k K
// rho K^dagger in ONE HBM pass: row-side and column-side passes (or a single
// superoperator pass) run on the tile while it is resident in shared memory.
#pragma once

#include <cuda.h>

#include "pw_internal.cuh"
#include "pw_blocks.cuh"

namespace pw {

constexpr int kTileThreads = 256;
constexpr int kMaxRuns = 5;

// One sparse-operator pass over a resident tile.
struct PassKind {
  uint32_t t;         // operator dimension
  uint32_t nfib;      // fibers of this pass inside one tile
  uint32_t nfib_pad;  // nfib rounded up to a multiple of 32 (warp-uniform rows)
  int nt;             // target axes, operator (given) order
  uint32_t tdim[PW_MAX_TARGETS * 2];
  uint32_t tstride[PW_MAX_TARGETS * 2];  // element strides inside the tile
  int nfr;                               // tile runs that are NOT targets of this pass
  uint32_t fr_size[kMaxRuns];            // outermost first
  uint32_t fr_stride[kMaxRuns];
};

struct TiledGeom {
  int rank;                     // tensor-map rank (= number of runs)
  uint64_t gdim[kMaxRuns];      // innermost first; dim 0 counted in REALS (2 per complex)
  uint64_t gstride[kMaxRuns];   // bytes, for dims 1..rank-1 (index 0 unused)
  uint32_t box[kMaxRuns];       // innermost first; box[0] in reals
  int nrest;                    // non-target runs, outermost first:
  uint32_t rest_box[kMaxRuns];      //   box size (complex elements)
  uint32_t rest_ntiles[kMaxRuns];   //   tiles along the run
  int rest_tmdim[kMaxRuns];         //   tensor-map dim index of the run
  uint64_t ntiles;
  uint32_t tile_elems;          // complex elements per tile
  int nkinds;
  PassKind kind[2];
};

// mode of the pass program executed per tile
enum { kModeVector = 0, kModeMatrixTwoSided = 1, kModeMatrixSuper = 2, kModeVectorConj = 3 };

struct TiledParams {
  TiledGeom g;
  int mode;
  int K;                 // operators (Kraus terms); 1 for a plain apply
  int nstage;            // input tiles in flight per CTA
  uint32_t csr_cap;      // entries reserved per operator
  const int* skip;       // optional device flag: exit when *skip != 0
  const int32_t* rowptr; // [K][t+1]
  const void* ents;      // [K][csr_cap] Entry<V>
  // block-structured form of the operator per pass kind (tile offsets); hdr == nullptr
  // or hdr->usable == 0 selects the CSR compute path instead
  BlockArrays blk[2];
};

template <typename V> struct Entry;
template <>
struct __align__(16) Entry<double2> {
  double2 val;
  int32_t off[2];  // tile offset of the column under pass kind 0 / 1
  int32_t pad[2];
};
template <>
struct __align__(16) Entry<float2> {
  float2 val;
  int32_t off[2];
};

// Plans the tile geometry.  `axes` is the full axis list of the tensor (for a density
// matrix: dims ++ dims), kind0/kind1 the target axes of the two pass kinds in
// operator order (nt1 = 0 for vectors), `boundary` an axis index at which runs
// must not merge (-1: none).  Returns false when the shape is not TMA-eligible.
bool plan_tiled(const int64_t* axes, int naxes, const int* kind0, int nt0, const int* kind1,
                int nt1, int boundary, size_t elem_bytes, uint32_t want_elems, TiledGeom* g);

// Launches CSR preparation + the tiled kernel.  ops: device (K, t, t).  mode as above.
int launch_tiled(const void* in, void* out, const void* ops, int K, const TiledGeom& g, int mode,
                 int dtype, cudaStream_t stream, const int* skip = nullptr);

// S (t^2 x t^2) = sum_k K_k (x) conj(K_k), row-major, on the device.
int build_superoperator(const void* ops, int K, uint32_t t, void* S, int dtype, cudaStream_t stream);

}  // namespace pw
