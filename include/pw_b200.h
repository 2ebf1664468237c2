/*
 * pw_b200.h -- C ABI of libpwb200.so: the B200 (sm_100a) drop-in for the
 * array-only seam of tqsd/photon_weave (photon_weave/core/adapters.py).
 *
 * The reference has no FFI of its own (it is pure Python on JAX); the seam a
 * replacement must bind is the function set of photon_weave/core/adapters.py
 * (SURVEY.md section 8b).  Each entry point below names the reference function
 * (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no framework types.
 *  - All state/operator pointers are DEVICE pointers unless the function name
 *    ends in _host.  Every device entry point only ENQUEUES work on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream): it never
 *    synchronises and never throws.  Return value: PW_OK or a PW_ERR_* code.
 *  - Layout: flat row-major tensor, subsystem 0 most significant; a state vector
 *    has prod(dims) elements, a density matrix prod(dims)^2 (row index = ket).
 *    Elements are interleaved complex: PW_C64 = 2 x float, PW_C128 = 2 x double.
 *  - `targets` are subsystem indices in the ORDER the operator's indices are
 *    factored (photon_weave/core/adapters.py:53 `perm = list(target_indices)+rest`);
 *    the tensor order of the state is never changed by apply/Kraus.
 *  - Operators are dense row-major (t x t), t = prod(dims[targets]), in the SAME
 *    dtype as the state (the Python host promotes, as JAX does).
 *  - `out` may alias the input state for pw_apply_vec / pw_apply_mat /
 *    pw_kraus_mat (in-place update); otherwise buffers must not overlap.
 *  - Probabilities are real arrays in the state's real dtype (float for PW_C64,
 *    double for PW_C128), normalised to sum 1, indexed in target-given order.
 */
#ifndef PW_B200_H
#define PW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PW_B200_VERSION 100 /* major*10000 + minor*100 + patch */

enum { PW_C64 = 0, PW_C128 = 1 };

enum {
  PW_OK = 0,
  PW_ERR_INVALID = 1,     /* bad dims/targets/null pointer                      */
  PW_ERR_UNSUPPORTED = 2, /* shape outside what the kernels cover               */
  PW_ERR_CUDA = 3,        /* a CUDA runtime call failed (pw_last_cuda_error())  */
  PW_ERR_WORKSPACE = 4    /* stream-ordered workspace allocation failed         */
};

#define PW_MAX_SUBSYSTEMS 32
#define PW_MAX_TARGETS 8

int pw_version(void);
const char* pw_status_string(int status);
/* cudaError_t of the last failing CUDA call on this host thread (0 if none). */
int pw_last_cuda_error(void);
/* Number of kernels this library has launched since load (bench "gpu_launches"). */
uint64_t pw_launch_count(void);
/* Apply paths enqueued since load: out[0] register kernel (t<=8), out[1] generic
 * fallback kernel, out[2] TMA-tiled kernel, out[3] tiled plans rejected (fell back),
 * out[4] block-structured attempts (structure analysed on the device; the fallback
 * enqueued behind it exits when the block kernel handled the operator). */
void pw_path_counters(uint64_t out[5]);
/* Tuning knobs (process-wide; also read once from the environment as PW_<NAME>):
 *   "vec_path"   0 auto | 1 register/generic kernels only | 2 TMA-tiled whenever eligible
 *                | 3 block-structured attempt for every t (then tiled / generic)
 *   "mat_mode"   0 auto | 1 two-pass fallback | 2 tiled two-sided | 3 tiled superoperator
 *                | 4 superoperator blocks, then two-pass (never the tiled kernel)
 *   "blocks_min_elems" states smaller than this skip the block-structure analysis
 *   "tile_elems" complex elements per shared-memory tile (0 = default)
 *   "ctas_per_sm" resident CTAs per SM for the tiled kernel (0 = derive from smem)
 * Returns PW_ERR_INVALID for an unknown name. */
int pw_set_option(const char* name, int value);

/* Arithmetic-peak probe (measurement aid: the FP64 roof the dense-operator kernels are quoted
 * against).  kind 0: FP64 FMA, 1: FP64 tensor cores (mma.sync.m8n8k4),
 * 2: FP32 FMA, 3: TF32 mma.sync.m16n8k8, 4..6: FP64 tensor cores with 1 / 2 / 3 FP64 adds interleaved per 6
 * MMAs, 7: with 2 FP32 adds (the cost of the three-product complex multiply; MMA flops only are counted).
 * Enqueues ONE register-only kernel of `iters` iterations; *flops (host) receives
 * the real flops it performs; time it with CUDA events.  `sink` is one device double. */
int pw_probe_arith(int kind, int iters, double* sink, double* flops, void* stream);

/* ---- analysed-plan cache ----------------------------------------------------------
 * pw_apply_vec / pw_apply_mat / pw_kraus_mat / pw_kraus_axes analyse their operator on the device
 * before the kernel that does the work (block structure, super-block packing, superoperator).
 * pw_op_pin(op) is the caller's promise that the device memory at `op` stays alive and UNCHANGED
 * until pw_op_unpin(op): the library then keeps the analysis per (operator, dims, targets, dtype,
 * aliasing) and later calls skip it (the container layer pins the cached device operator of an
 * Operation -- photon_weave/operation/operation.py:292-294 recomputes it on every access).
 * Results are identical with and without pinning.  pw_op_unpin releases the cached buffers (it
 * waits for the device).  pw_plan_cache_size: entries alive (tests). */
int pw_op_pin(const void* op);
int pw_op_unpin(const void* op);
int pw_plan_cache_size(void);

/* ---- apply ------------------------------------------------------------------ */

/* psi' = (O on targets) psi.
 * Replaces kernels.apply_op_vector_ordered (core/kernels.py:62-98) and the
 * reorder/einsum/restore around it in adapters.apply_operation_vector
 * (core/adapters.py:299-345), both use_contraction routes. */
int pw_apply_vec(const void* psi, void* out, const void* op, const int64_t* dims, int n,
                 const int32_t* targets, int nt, int dtype, void* stream);

/* rho' = (O x I) rho (O x I)^dagger.
 * Replaces kernels.apply_op_matrix_ordered (core/kernels.py:101-145) and
 * adapters.apply_operation_matrix (core/adapters.py:348-393). */
int pw_apply_mat(const void* rho, void* out, const void* op, const int64_t* dims, int n,
                 const int32_t* targets, int nt, int dtype, void* stream);

/* rho' = sum_k (K_k x I) rho (K_k x I)^dagger, ops stacked (K, t, t).
 * Replaces kernels.apply_kraus_matrix_ordered (core/kernels.py:148-186),
 * adapters.apply_kraus_matrix (core/adapters.py:396-466) and the un-embedded
 * _math/ops.apply_kraus (_math/ops.py:371-403, dims = (d,)). */
int pw_kraus_mat(const void* rho, void* out, const void* ops, int K, const int64_t* dims,
                 int n, const int32_t* targets, int nt, int dtype, void* stream);

/* Two single-subsystem operators on DIFFERENT axes of equal dimension d (2..6) in one HBM
 * pass: psi' = (A on axis_a)(B on axis_b) psi (the two commute).  Half the traffic of two
 * pw_apply_vec calls.  PW_ERR_UNSUPPORTED outside that shape. */
int pw_apply_vec_pair(const void* psi, void* out, const void* op_a, const void* op_b,
                      const int64_t* dims, int n, int axis_a, int axis_b, int dtype, void* stream);

/* `nops` (2..4) single-subsystem operators on DIFFERENT axes of one common dimension D (2..8) in
 * ONE HBM pass: psi' = prod_i (ops[i] on axes[i]) psi (they commute).  `ops` is a HOST array of
 * device pointers.  The sub-tensor spanned by the axes stays resident in shared memory while the
 * operators are applied one after the other, so a layer of single-mode operators
 * (fock.py:427 -> composite_envelope.py:587 once per mode) crosses HBM once per nops operators.
 * `out` may alias `psi`.  PW_ERR_UNSUPPORTED outside that shape (apply them one by one). */
int pw_apply_vec_multi(const void* psi, void* out, const void* const* ops, const int64_t* dims, int n,
                       const int32_t* axes, int nops, int dtype, void* stream);

/* Fused multi-step evolution of a small state (prod(dims) <= 128) under ONE dense operator
 * acting on ALL subsystems (op is N x N, indices in tensor order): psi <- op^nsteps psi in a
 * single kernel.  If reduced_out != NULL it receives, for every step s, the reduced density
 * matrix (tk x tk, tk = prod(dims[keep]) <= 8) of the state AFTER step s, i.e. what
 * `apply_operation` + `trace_out` return per iteration of the Jaynes-Cummings loop
 * (examples/jaynes_cummings_model.py:130-141).  New surface: the reference has no
 * multi-step entry point. */
int pw_evolve_vec(const void* psi, void* out, const void* op, const int64_t* dims, int n,
                  const int32_t* keep, int nkeep, int nsteps, void* reduced_out, int dtype,
                  void* stream);

/* ---- reductions ---------------------------------------------------------------- */

/* out (t x t) = Tr_rest rho, kept subsystems ordered as listed in `keep`.
 * Replaces kernels.trace_out_matrix_ordered (core/kernels.py:301-319) and
 * adapters.trace_out_matrix (core/adapters.py:696-733): pass `keep` as given for
 * the ordered route (:720,732), sorted ascending for the contraction route (:80-81). */
int pw_trace_out_mat(const void* rho, void* out, const int64_t* dims, int n,
                     const int32_t* keep, int nkeep, int dtype, void* stream);

/* out (t x t) = Tr_rest |psi><psi| without materialising psi psi^dagger.
 * Replaces state/utils/trace_out.trace_out_vector (state/utils/trace_out.py:17-25)
 * and the second output of kernels.measure_vector_expectation_ordered
 * (core/kernels.py:266-283) when `keep` is the unmeasured subsystems. */
int pw_reduced_from_vec(const void* psi, void* out, const int64_t* dims, int n,
                        const int32_t* keep, int nkeep, int dtype, void* stream);

/* out (t x t)[a, b] = sum_rest x[(a, rest)] conj(y[(b, rest)]) over a tensor of `naxes` axes (up
 * to 64: pass dims ++ dims for density matrices), kept axes in the given order.  With x = the
 * cotangent of the output and y = the input state this is the OPERATOR cotangent of the linear
 * maps above: the transform rule the reference gets from jax.grad through its einsum
 * (core/kernels.py:87-98; tests/state/test_envelope.py:846-870) -- SURVEY.md section 8f.4. */
int pw_cross_reduced(const void* x, const void* y, void* out, const int64_t* axes, int naxes,
                     const int32_t* keep, int nkeep, int dtype, void* stream);

/* probs[j] = sum_rest |psi[j,rest]|^2 / sum_all.
 * Replaces kernels._measurement_probs_vector (core/kernels.py:28-41). */
int pw_probs_vec(const void* psi, void* probs, const int64_t* dims, int n,
                 const int32_t* targets, int nt, int dtype, void* stream);

/* probs[j] = real(sum_rest rho[(j,rest),(j,rest)]) / sum_all (reads only the diagonal).
 * Replaces kernels._measurement_probs_matrix (core/kernels.py:44-59). */
int pw_probs_mat(const void* rho, void* probs, const int64_t* dims, int n,
                 const int32_t* targets, int nt, int dtype, void* stream);

/* ---- sampling + collapse ---------------------------------------------------- */

/* *outcome = jax.random.choice(key, a=t, p=probs): threefry2x32 bits of the
 * scalar draw, cumsum, searchsorted-left; computed on the device, no host sync.
 * Replaces the jax.random.choice call sites core/kernels.py:204,224,243,261,341. */
int pw_draw(const void* probs, int64_t t, uint32_t key0, uint32_t key1, int dtype,
            int64_t* outcome, void* stream);

/* Same draw with the key's two uint32 words in DEVICE memory (a traced jax.random key inside a
 * jitted computation: the XLA FFI handler cannot read it on the host without a sync). */
int pw_draw_devkey(const void* probs, int64_t t, const uint32_t* key_dev, int dtype,
                   int64_t* outcome, void* stream);

/* post (r) = psi[outcome, :] / sqrt(probs[outcome] + 1e-18); rest in tensor order.
 * `outcome` and `probs` are device pointers (outputs of pw_draw / pw_probs_vec).
 * probs == NULL: plain gather without rescaling (jnp.take of the legacy sequential path,
 * state/utils/measurements.py:135).  Replaces core/kernels.py:205 (and :244). */
int pw_collapse_vec(const void* psi, void* post, const int64_t* outcome, const void* probs,
                    const int64_t* dims, int n, const int32_t* targets, int nt, int dtype,
                    void* stream);

/* post (r x r) = rho[outcome, :, outcome, :] / (probs[outcome] + 1e-18).
 * Replaces core/kernels.py:225-226 (and :262). */
int pw_collapse_mat(const void* rho, void* post, const int64_t* outcome, const void* probs,
                    const int64_t* dims, int n, const int32_t* targets, int nt, int dtype,
                    void* stream);

/* POVM probabilities from the reduced state of the targets:
 * probs[k] = real Tr(E_k rho_T E_k^dagger) / sum, traces[k] = un-normalised value.
 * rho_T is (t x t) (pw_trace_out_mat with keep = targets), ops (K, t, t).
 * Replaces the K full-size projections + traces of
 * kernels.measure_povm_matrix_ordered (core/kernels.py:331-339). */
int pw_povm_probs(const void* rho_t, const void* ops, int K, int64_t t, void* probs,
                  void* traces, int dtype, void* stream);

/* op_out (t x t) = ops[*outcome] / sqrt(traces[*outcome] + 1e-18), so that
 * pw_apply_mat(rho, op_out) is the POVM post state of core/kernels.py:342-343. */
int pw_povm_select(const void* ops, int64_t t, const int64_t* outcome, const void* traces,
                   void* op_out, int dtype, void* stream);

/* ---- bookkeeping reductions the callers need ------- */

/* stats[0] = sum |x|^2, stats[1] = max |x|^2 over `count` complex elements (doubles). */
int pw_norm_stats(const void* x, int64_t count, int dtype, double* stats, void* stream);

/* Fused post-apply bookkeeping, ONE pass over psi (is_matrix = 0) or over the diagonal of
 * rho (is_matrix = 1):  levels[i] = highest basis index of subsystem i that carries a
 * non-zero amplitude / population (-1 when the state is all zeros), for every subsystem at
 * once; stats[0] = sum |psi|^2 (or Tr rho), stats[1] = max |x|^2.  levels is n device
 * int32, stats two device doubles; n <= 16.  Replaces the per-Fock-mode
 * trace_out -> psi psi^dagger -> num_quanta_{vector,matrix} loop the reference runs after
 * every composite apply (state/composite_envelope.py:680-696, _math/ops.py:472-512); equal
 * to it for physical (positive semi-definite) states. */
int pw_occupancy(const void* state, const int64_t* dims, int n, int is_matrix, int dtype,
                 int32_t* levels, double* stats, void* stream);

/* x *= *scale_dev (a device double); e.g. 1/sqrt(norm) for Operation.renormalize
 * (state/composite_envelope.py:685-686). */
int pw_scale(void* x, int64_t count, const double* scale_dev, int dtype, void* stream);

/* x /= *divisor_dev (true division): `state / jnp.linalg.norm(state)` with the reference's
 * rounding, so a renormalised basis state holds exactly 1.0
 * (state/fock.py:493-495, state/composite_envelope.py:685-686). */
int pw_divide(void* x, int64_t count, const double* divisor_dev, int dtype, void* stream);

/* ---- layout ------------------------------------------------------------------ */

/* out = transpose(in.reshape(dims), perm), contiguous (out axis k = in axis perm[k]).
 * Replaces the jnp.transpose passes of reorder_vector_front / reorder_matrix_front
 * (core/adapters.py:210-296) and ProductState.reorder
 * (state/composite_envelope.py:201-248; pass dims ++ dims for a density matrix);
 * also the pack step of the multi-GPU axis swap.  in and out must not alias. */
int pw_permute(const void* in, void* out, const int64_t* dims, int n, const int32_t* perm,
               int dtype, void* stream);

/* out = the same tensor with dims[axis] = new_dim (zero padding or truncation); for a density
 * matrix (is_matrix) both the row and the column copy of the subsystem are resized.
 * Replaces ProductState.resize_fock (state/composite_envelope.py:495-585; the caller does the
 * occupancy check with pw_probs_* before shrinking). */
int pw_resize_axis(const void* in, void* out, const int64_t* dims, int n, int axis,
                   int64_t new_dim, int is_matrix, int dtype, void* stream);

/* out = kron(a, b) for row-major (rows x cols) operands; column vectors have cols = 1.
 * Replaces kron_reduce in CompositeEnvelope.combine (core/ops.py:53-64,
 * state/composite_envelope.py:915-948). */
int pw_kron(const void* a, const void* b, void* out, int64_t rows_a, int64_t cols_a,
            int64_t rows_b, int64_t cols_b, int dtype, void* stream);

/* out[0] = how many entries of psi equal exactly 1 + 0i, out[1] = index of the first one
 * (-1 if none); out is two device int64.  The Vector -> Label test of state_contract
 * (state/utils/state_transform.py:138-144). */
int pw_find_label(const void* psi, int64_t n, int dtype, int64_t* out, void* stream);

/* rho = psi psi^dagger (state_expand Vector -> Matrix, state/utils/state_transform.py:58-63). */
int pw_expand_vec(const void* psi, void* rho, int64_t n, int dtype, void* stream);

/* state_contract Matrix -> Vector (state/utils/state_transform.py:122-137) without rho@rho and
 * eigh: info[0] = Tr(rho^2) = sum |rho_ij|^2, info[1] = Tr rho; *flag = |Tr(rho^2) - 1| < tol;
 * when set, psi = the pure state with the global phase fixed so that psi[0] is real >= 0
 * (the reference's convention).  info / flag are device pointers. */
int pw_contract_mat(const void* rho, int64_t n, double tol, void* psi, double* info, int* flag,
                    int dtype, void* stream);

/* ---- multi-GPU: peer memory over NVLink (one process per GPU) ---------------------- */

/* cudaMalloc + CUDA IPC export of a shard buffer; peers map it with pw_peer_open. */
int pw_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int pw_peer_open(const unsigned char handle[64], void** ptr);
int pw_peer_close(void* ptr);
int pw_peer_free(void* ptr);

/* dst (local) <- src (local or IPC-mapped peer memory), `bytes` bytes, on a COPY ENGINE
 * (cudaMemcpyAsync): one piece of the pipelined axis swap, overlapping the apply kernels. */
int pw_peer_copy(void* dst, const void* src, size_t bytes, void* stream);

/* Axis swap as a direct peer read: out[p * chunk + x] = srcs[p][x] for p < W.
 * `srcs` is a HOST array of W device pointers (local or IPC-mapped peer memory), each
 * already offset to the chunk this rank needs.  `rank` rotates the traversal so the W
 * ranks pull from W different peers at any moment.  No collective library call. */
int pw_gather_chunks(const void* const* srcs, int W, int64_t chunk_elems, int rank, void* out,
                     int dtype, void* stream);

/* The same swap FUSED with the operator application that follows it: the input of
 * pw_apply_vec is the virtual concatenation of the W peer chunks (dims describe that
 * concatenated tensor, prod(dims) = W * chunk_elems), loads go over NVLink, the result
 * is written once to local memory.  Covers t <= 8 and, for 8 < t <= 128, BLOCK-STRUCTURED
 * operators (beam splitters, two-mode phases: the device-side structure analysis of
 * pw_apply_vec; its verdict is read back, so this one entry point synchronises the stream);
 * PW_ERR_UNSUPPORTED otherwise (dense t > 8: gather, then pw_apply_vec). */
int pw_apply_vec_gather(const void* const* srcs, int W, int64_t chunk_elems, int rank, void* out,
                        const void* op, const int64_t* dims, int n, const int32_t* targets,
                        int nt, int dtype, void* stream);

/* ---- multi-GPU layer: blocks of a sharded state -------------------------------------------
 * One process per GPU holds a BLOCK of the state (photon_weave_b200/sharded.py): a row slab of
 * rho, or a slab of psi over virtual axes.  Such a block is not a dense tensor of the global
 * dims, so these entry points take explicit axes / (extent, element stride) lists.  The
 * reference has no distributed code; the arithmetic per block is that of the functions cited. */

/* out = sum_k (K_k on row_targets) x (conj(K_k) on col_targets), x a row-major tensor over
 * `naxes` axes (a rectangular block of rho: row axes then column axes).  K = 1: U rho U^dagger.
 * Same kernels and one-pass dispatch as pw_apply_mat / pw_kraus_mat (core/kernels.py:101-186),
 * which are the special case axes = dims ++ dims. */
int pw_kraus_axes(const void* x, void* out, const void* ops, int K, const int64_t* axes, int naxes,
                  const int32_t* row_targets, int ntr, const int32_t* col_targets, int ntc,
                  int dtype, void* stream);

/* sums[j] = sum over the `rest` axes of |x|^2 (real_part = 0) or Re x (real_part = 1) at target
 * index j (mixed radix over the tgt axes, first most significant): the UN-NORMALISED partial
 * sums of core/kernels.py:28-59 on one rank's block; `x` is already offset to the block (and,
 * for rho, strides walk its diagonal).  sums is t device doubles. */
int pw_sums_strided(const void* x, const int64_t* rest_extent, const int64_t* rest_stride, int nrest,
                    const int64_t* tgt_extent, const int64_t* tgt_stride, int ntgt, int real_part,
                    double* sums, int dtype, void* stream);

/* probs[j] = sums[j] / sum(sums) in the state's real dtype (after the cross-rank all-reduce). */
int pw_probs_from_sums(const double* sums, int64_t t, void* probs, int dtype, void* stream);

/* out[j, j'] = sum over the rest axes of x[rest + row(j) + col(j')]: partial trace of a block of
 * rho (core/kernels.py:301-319) whose traced index pairs are walked by the rest strides. */
int pw_trace_strided(const void* x, const int64_t* rest_extent, const int64_t* rest_stride, int nrest,
                     const int64_t* row_extent, const int64_t* row_stride, int nrow,
                     const int64_t* col_extent, const int64_t* col_stride, int ncol, void* out,
                     int dtype, void* stream);

/* out (contiguous, row-major over the axes) = x[strided view] / den; mode 0: den = 1,
 * 1: sqrt(probs[*outcome] + 1e-18) (core/kernels.py:205), 2: probs[*outcome] + 1e-18 (:225). */
int pw_gather_strided(const void* x, const int64_t* extent, const int64_t* stride, int naxes,
                      void* out, const void* probs, const int64_t* outcome, int mode, int dtype,
                      void* stream);

/* ---- host-side PRNG key management (pure integer, no device) --------------- */

/* out[i] = jax.random.split(key, num)[i] (partitionable threefry2x32).
 * Replaces jax.random.split at core/rng.py:37, photon_weave.py:47. */
int pw_rng_split(const uint32_t key[2], int num, uint32_t* out /* num x 2 */);

/* ---- host-buffer entry points (end-to-end path: H2D + kernels + D2H) -------- */

typedef struct {
  int kind;          /* 0 = operator on a state vector, 1 = operator on a density
                        matrix, 2 = Kraus channel on a density matrix              */
  int nt;            /* number of target subsystems                               */
  int32_t targets[PW_MAX_TARGETS];
  int K;             /* number of Kraus operators (kind 2), else 1                */
  const void* op;    /* HOST pointer, (K,) t x t row-major, state dtype           */
} pw_host_op;

/* Upload h_state (pageable or pinned host memory), run `nops` operations in
 * order on the device, download into h_out (may alias h_state).  is_matrix
 * selects vector / density-matrix layout.  Synchronises before returning. */
int pw_circuit_host(const void* h_state, void* h_out, int is_matrix, const int64_t* dims,
                    int n, const pw_host_op* ops, int nops, int dtype, int device);

/* Device-resident operation list: pw_host_op.op are DEVICE pointers here.  Ping-pongs
 * between `state` and `scratch` (scratch may be NULL: in place); *result_in_scratch tells
 * which buffer holds the result.  fuse != 0: adjacent single-subsystem vector operators on
 * different axes (they commute) are applied as one pass (pw_apply_vec_pair).  Enqueue only. */
int pw_circuit_device(void* state, void* scratch, int is_matrix, const int64_t* dims, int n,
                      const pw_host_op* ops, int nops, int dtype, int fuse, void* stream,
                      int* result_in_scratch);

/* ---- streaming host-buffer pipeline ------------------------------------------------
 * Throughput version of pw_circuit_host for callers that evolve many states (sweeps,
 * trajectories): up to `nbuf` (2..8, 4 sustains the PCIe rate) state buffers stay in HBM and
 * the upload of job k+1, the kernels of job k and the download of job k-1 run concurrently on
 * three streams, so PCIe is used full duplex.  submit only enqueues (h_state / h_out should be
 * pinned and must stay valid until the job's wait returns; they may not alias another job in
 * flight); wait blocks until that job's result is in h_out.  Not thread safe per pipeline. */
typedef struct pw_pipeline pw_pipeline;
int pw_pipeline_create(int is_matrix, const int64_t* dims, int n, int dtype, int device,
                       int nbuf, pw_pipeline** out);
int pw_pipeline_buffers(const pw_pipeline* pl);   /* buffers actually allocated (>= 2) */
int pw_pipeline_submit(pw_pipeline* pl, const void* h_state, void* h_out, const pw_host_op* ops,
                       int nops, int64_t* ticket);
int pw_pipeline_wait(pw_pipeline* pl, int64_t ticket);
int pw_pipeline_destroy(pw_pipeline* pl);

#ifdef __cplusplus
}
#endif
#endif /* PW_B200_H */
