// Host-buffer entry point: the end-to-end path a caller with NumPy/host arrays
// takes (H2D upload, the operation list on the device, D2H download).
#include <vector>

#include "pw_internal.cuh"

using namespace pw;

namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};
struct Stream {
  cudaStream_t s = nullptr;
  ~Stream() { if (s) cudaStreamDestroy(s); }
};
}  // namespace

// Runs an operation list on device buffers, ping-ponging between `cur` and `other`
// (other == nullptr: in place).  Adjacent single-subsystem vector operators on different
// axes commute; with `fuse` they are applied as ONE pass (pw_apply_vec_pair).
static int run_ops(void** cur, void** other, int is_matrix, const int64_t* dims, int n,
                   const pw_host_op* ops, void* const* dev_ops, int nops, int dtype, int fuse,
                   cudaStream_t s) {
  int i = 0;
  while (i < nops) {
    const pw_host_op& o = ops[i];
    void* dst = *other ? *other : *cur;
    int rc = PW_ERR_UNSUPPORTED;
    int used = 1;
    if (fuse && o.kind == 0 && o.nt == 1) {
      // consecutive single-subsystem operators on DIFFERENT axes commute: up to `fuse_group` of
      // one common dimension go through the resident-tile kernel in ONE pass
      const int gmax = pw::options().fuse_group.load(std::memory_order_relaxed);
      int cnt = 1;
      int32_t axes[4] = {o.targets[0], 0, 0, 0};
      const void* gops[4] = {dev_ops[i], nullptr, nullptr, nullptr};
      while (cnt < gmax && cnt < 4 && i + cnt < nops && ops[i + cnt].kind == 0 && ops[i + cnt].nt == 1 &&
             dims[ops[i + cnt].targets[0]] == dims[o.targets[0]]) {
        bool dup = false;
        for (int k = 0; k < cnt; ++k) dup = dup || axes[k] == ops[i + cnt].targets[0];
        if (dup) break;
        axes[cnt] = ops[i + cnt].targets[0];
        gops[cnt] = dev_ops[i + cnt];
        ++cnt;
      }
      // measured on the 34.8 GB vector (profiles/r2_multi_*.txt): three or four operators per
      // pass win everywhere in complex128 (1.28-1.47x); pairs go to the register-tile pair kernel
      // below; complex64 tiles whose contiguous direction is a target lose (0.7x) and stay unfused
      bool inner = false;
      for (int k = 0; k < cnt; ++k) inner = inner || axes[k] == n - 1;
      if (dtype == PW_C64 && inner) cnt = 0;
      while (cnt >= 3 && rc == PW_ERR_UNSUPPORTED) {
        rc = pw_apply_vec_multi(*cur, dst, gops, dims, n, axes, cnt, dtype, s);
        if (rc == PW_OK) used = cnt; else if (rc == PW_ERR_UNSUPPORTED) --cnt;
      }
    }
    if (rc == PW_ERR_UNSUPPORTED && fuse && *other && o.kind == 0 && o.nt == 1 && i + 1 < nops &&
        ops[i + 1].kind == 0 && ops[i + 1].nt == 1 && ops[i + 1].targets[0] != o.targets[0] &&
        // complex64 with the contiguous axis among the targets: the pair kernel's 8-byte strided
        // accesses lose to two coalesced passes (21.4 against 7.6 + 7.0 ms on the 17.4 GB vector)
        !(dtype == PW_C64 && (o.targets[0] == n - 1 || ops[i + 1].targets[0] == n - 1))) {
      rc = pw_apply_vec_pair(*cur, dst, dev_ops[i], dev_ops[i + 1], dims, n, o.targets[0],
                             ops[i + 1].targets[0], dtype, s);
      if (rc == PW_OK) used = 2;
    }
    if (rc == PW_ERR_UNSUPPORTED) {
      if (o.kind == 0)
        rc = pw_apply_vec(*cur, dst, dev_ops[i], dims, n, o.targets, o.nt, dtype, s);
      else if (o.kind == 1)
        rc = pw_apply_mat(*cur, dst, dev_ops[i], dims, n, o.targets, o.nt, dtype, s);
      else
        rc = pw_kraus_mat(*cur, dst, dev_ops[i], o.K, dims, n, o.targets, o.nt, dtype, s);
    }
    if (rc != PW_OK) return rc;
    if (*other) { void* t = *cur; *cur = *other; *other = t; }
    i += used;
  }
  (void)is_matrix;
  return PW_OK;
}

// bytes of the state (vector: N, density matrix: N^2 elements) with the overflow guard of
// build_generic (pw_plan.cu): anything beyond 2^62 elements is rejected, never wrapped.
static int state_bytes(int is_matrix, const int64_t* dims, int n, size_t es, size_t* bytes) {
  uint64_t N = 1;
  for (int i = 0; i < n; ++i) {
    if (dims[i] < 1) return PW_ERR_INVALID;
    if (N > (1ull << 62) / (uint64_t)dims[i]) return PW_ERR_INVALID;
    N *= (uint64_t)dims[i];
  }
  uint64_t count = N;
  if (is_matrix) {
    if (N > (1ull << 31)) return PW_ERR_INVALID;  // N^2 * 16 must stay below 2^66 -> cap at 2^62 elements
    count = N * N;
  }
  if (count > (1ull << 62) / es) return PW_ERR_INVALID;
  *bytes = (size_t)(count * es);
  return PW_OK;
}

static int check_ops(int is_matrix, const int64_t* dims, int n, const pw_host_op* ops, int nops) {
  for (int i = 0; i < nops; ++i) {
    const pw_host_op& o = ops[i];
    if (o.nt < 1 || o.nt > PW_MAX_TARGETS || !o.op) return PW_ERR_INVALID;
    if ((o.kind != 0) != (is_matrix != 0)) return PW_ERR_INVALID;
    if (o.kind == 2 && o.K < 1) return PW_ERR_INVALID;
    uint32_t seen = 0;
    uint64_t t = 1;
    for (int k = 0; k < o.nt; ++k) {
      const int a = o.targets[k];
      if (a < 0 || a >= n || (seen >> a & 1u)) return PW_ERR_INVALID;  // out of range / duplicate
      seen |= 1u << a;
      t *= (uint64_t)dims[a];
      if (t > 0xffffffffull) return PW_ERR_UNSUPPORTED;
    }
  }
  return PW_OK;
}

extern "C" int pw_circuit_device(void* state, void* scratch, int is_matrix, const int64_t* dims,
                                 int n, const pw_host_op* ops, int nops, int dtype, int fuse,
                                 void* stream, int* result_in_scratch) {
  if (!state || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || nops < 0 || (nops && !ops)) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  PW_TRY(check_ops(is_matrix, dims, n, ops, nops));
  std::vector<void*> dev_ops(nops);
  for (int i = 0; i < nops; ++i) dev_ops[i] = const_cast<void*>(ops[i].op);
  void* cur = state;
  void* other = scratch;
  PW_TRY(run_ops(&cur, &other, is_matrix, dims, n, ops, dev_ops.data(), nops, dtype, fuse,
                 (cudaStream_t)stream));
  if (result_in_scratch) *result_in_scratch = (cur != state) ? 1 : 0;
  return PW_OK;
}

extern "C" int pw_circuit_host(const void* h_state, void* h_out, int is_matrix,
                               const int64_t* dims, int n, const pw_host_op* ops, int nops,
                               int dtype, int device) {
  if (!h_state || !h_out || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || nops < 0 || (nops && !ops))
    return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  const size_t es = dtype == PW_C128 ? 16 : 8;
  size_t bytes = 0;
  PW_TRY(state_bytes(is_matrix, dims, n, es, &bytes));
  PW_TRY(check_ops(is_matrix, dims, n, ops, nops));
  PW_CUDA(cudaSetDevice(device));

  Stream st;
  PW_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
  DevBuf state, other;
  PW_CUDA(cudaMalloc(&state.p, bytes));
  // ping-pong when a second buffer fits (lets the block-structured path run); else in place
  if (cudaMalloc(&other.p, bytes) != cudaSuccess) {
    other.p = nullptr;
    (void)cudaGetLastError();
  }

  // operators: one small upload each, queued ahead of the state so they overlap it
  std::vector<DevBuf> dops(nops);
  for (int i = 0; i < nops; ++i) {
    const pw_host_op& o = ops[i];  // validated by check_ops above
    uint64_t t = 1;
    for (int k = 0; k < o.nt; ++k) t *= (uint64_t)dims[o.targets[k]];
    const int K = o.kind == 2 ? o.K : 1;
    const size_t ob = (size_t)K * t * t * es;
    PW_CUDA(cudaMalloc(&dops[i].p, ob));
    PW_CUDA(cudaMemcpyAsync(dops[i].p, o.op, ob, cudaMemcpyHostToDevice, st.s));
  }
  std::vector<void*> dev_ops(nops);
  for (int i = 0; i < nops; ++i) dev_ops[i] = dops[i].p;
  const int fuse = pw::options().fuse_pairs.load(std::memory_order_relaxed);
  // ---- upload overlapped with compute ------------------------------------------------------
  // Operations on disjoint subsystems commute.  Those that do not touch subsystem 0 (nor follow,
  // on a shared subsystem, one that does) act independently on every slab of the leading axis, so
  // they run slab by slab BEHIND the upload (a second stream, one event per slab) instead of
  // waiting for the whole 34.8 GB state; the rest runs on the complete state afterwards.
  std::vector<int> late(nops, 0);
  int n_early = 0;
  const int64_t slabs = dims[0];
  const int overlap_mib = pw::options().host_overlap.load(std::memory_order_relaxed);  // 0 = off, else minimum state size
  if (!is_matrix && n >= 2 && slabs >= 2 && slabs <= 64 && overlap_mib > 0 && bytes >= ((size_t)overlap_mib << 20)) {
    for (int i = 0; i < nops; ++i) {
      for (int k = 0; k < ops[i].nt; ++k) late[i] |= ops[i].targets[k] == 0;
      for (int j = 0; j < i && !late[i]; ++j) {
        if (!late[j]) continue;
        for (int a = 0; a < ops[i].nt && !late[i]; ++a)
          for (int b = 0; b < ops[j].nt; ++b)
            if (ops[i].targets[a] == ops[j].targets[b]) { late[i] = 1; break; }
      }
      n_early += !late[i];
    }
  }
  if (n_early > 0) {
    Stream up;
    PW_CUDA(cudaStreamCreateWithFlags(&up.s, cudaStreamNonBlocking));
    std::vector<pw_host_op> early, rest;
    std::vector<void*> early_dev, rest_dev;
    for (int i = 0; i < nops; ++i) {
      if (late[i]) { rest.push_back(ops[i]); rest_dev.push_back(dev_ops[i]); continue; }
      pw_host_op o = ops[i];
      for (int k = 0; k < o.nt; ++k) o.targets[k] -= 1;   // the slab drops the leading axis
      early.push_back(o);
      early_dev.push_back(dev_ops[i]);
    }
    const size_t slab_bytes = bytes / (size_t)slabs;
    void* cur0 = state.p;
    void* oth0 = other.p;
    void* res_cur = state.p;
    void* res_oth = other.p;
    std::vector<cudaEvent_t> evs((size_t)slabs, nullptr);
    int rc = PW_OK;
    for (int64_t k = 0; k < slabs && rc == PW_OK; ++k) {
      const size_t off = (size_t)k * slab_bytes;
      rc = cuda_status(cudaMemcpyAsync((char*)cur0 + off, (const char*)h_state + off, slab_bytes,
                                       cudaMemcpyHostToDevice, up.s));
      if (rc == PW_OK) rc = cuda_status(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
      if (rc == PW_OK) rc = cuda_status(cudaEventRecord(evs[k], up.s));
      if (rc == PW_OK) rc = cuda_status(cudaStreamWaitEvent(st.s, evs[k], 0));
      if (rc != PW_OK) break;
      void* c = (char*)cur0 + off;
      void* o = oth0 ? (char*)oth0 + off : nullptr;
      rc = run_ops(&c, &o, 0, dims + 1, n - 1, early.data(), early_dev.data(), (int)early.size(), dtype, fuse, st.s);
      // every slab takes the same passes: the result sits in the same buffer for all of them
      res_cur = ((char*)c - off);
      res_oth = o ? (void*)((char*)o - off) : nullptr;
    }
    if (rc == PW_OK) {
      state.p = res_cur;
      other.p = res_oth;
      rc = run_ops(&state.p, &other.p, 0, dims, n, rest.data(), rest_dev.data(), (int)rest.size(), dtype, fuse, st.s);
    }
    if (rc != PW_OK) {
      cudaStreamSynchronize(up.s);
      cudaStreamSynchronize(st.s);
      for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
      return rc;
    }
    PW_CUDA(cudaMemcpyAsync(h_out, state.p, bytes, cudaMemcpyDeviceToHost, st.s));
    const int rs = cuda_status(cudaStreamSynchronize(st.s));
    cudaStreamSynchronize(up.s);
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    return rs;
  }
  PW_CUDA(cudaMemcpyAsync(state.p, h_state, bytes, cudaMemcpyHostToDevice, st.s));
  {
    const int rc = run_ops(&state.p, &other.p, is_matrix, dims, n, ops, dev_ops.data(), nops, dtype,
                           fuse, st.s);
    if (rc != PW_OK) {
      cudaStreamSynchronize(st.s);
      return rc;
    }
  }
  PW_CUDA(cudaMemcpyAsync(h_out, state.p, bytes, cudaMemcpyDeviceToHost, st.s));
  PW_CUDA(cudaStreamSynchronize(st.s));
  return PW_OK;
}

// ---------------------------------------------------------------------------------------
// Streaming host-buffer pipeline.
//
// pw_circuit_host is latency-bound by PCIe: upload, compute and download of one call run
// back to back (34.8 GB each way for the 12-mode vector = 2 x ~0.65 s around 0.22 s of
// kernels).  A caller that evolves many states (parameter sweeps, trajectories, batches)
// wants THROUGHPUT: this pipeline keeps up to `nbuf` state buffers in HBM and three
// streams (H2D, compute, D2H), so the upload of job k+1 and the download of job k-1 run
// on the two DMA engines while job k computes -- PCIe is used full duplex and the per-job
// time approaches max(upload, download).  Everything is stream ordered (events); submit
// never blocks on the device, wait blocks on one job's download only.
//
// Buffer pool: a job takes the two least recently released buffers (X for the upload, Z as
// the ping-pong scratch); after its kernels the buffer that does not hold the result is
// released at once, the result buffer when its download completes.  A buffer stays busy
// for upload + compute + download, so nbuf = 4 (3 jobs in flight + scratch) sustains the
// PCIe rate; nbuf = 2 degenerates to the serial schedule of pw_circuit_host.
// ---------------------------------------------------------------------------------------
#include <deque>

struct pw_pipeline {
  int device = 0, is_matrix = 0, n = 0, dtype = 0;
  int64_t dims[PW_MAX_SUBSYSTEMS];
  size_t bytes = 0, es = 0;
  cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr, s_ops = nullptr;
  struct Buf { void* p; cudaEvent_t free_ev; };
  std::deque<Buf> pool;            // least recently released first
  std::vector<void*> owned;        // for destruction
  struct Slot { void* arena = nullptr; size_t cap = 0; cudaEvent_t ops_ev = nullptr, done_ev = nullptr;
                int64_t ticket = -1; };
  std::vector<Slot> slots;         // ring of in-flight jobs (operator arenas + completion events)
  int64_t next_ticket = 0;
};

extern "C" int pw_pipeline_destroy(pw_pipeline* pl) {
  if (!pl) return PW_OK;
  cudaSetDevice(pl->device);
  if (pl->s_d2h) cudaStreamSynchronize(pl->s_d2h);
  if (pl->s_comp) cudaStreamSynchronize(pl->s_comp);
  if (pl->s_h2d) cudaStreamSynchronize(pl->s_h2d);
  for (auto& b : pl->pool) if (b.free_ev) cudaEventDestroy(b.free_ev);
  for (void* p : pl->owned) cudaFree(p);
  for (auto& s : pl->slots) {
    if (s.arena) cudaFree(s.arena);
    if (s.ops_ev) cudaEventDestroy(s.ops_ev);
    if (s.done_ev) cudaEventDestroy(s.done_ev);
  }
  for (cudaStream_t s : {pl->s_h2d, pl->s_comp, pl->s_d2h, pl->s_ops}) if (s) cudaStreamDestroy(s);
  delete pl;
  return PW_OK;
}

extern "C" int pw_pipeline_create(int is_matrix, const int64_t* dims, int n, int dtype, int device,
                                  int nbuf, pw_pipeline** out) {
  if (!out || !dims || n < 1 || n > PW_MAX_SUBSYSTEMS || nbuf < 2 || nbuf > 8) return PW_ERR_INVALID;
  if (dtype != PW_C64 && dtype != PW_C128) return PW_ERR_INVALID;
  PW_CUDA(cudaSetDevice(device));
  pw_pipeline* pl = new pw_pipeline();
  pl->device = device; pl->is_matrix = is_matrix; pl->n = n; pl->dtype = dtype;
  pl->es = dtype == PW_C128 ? 16 : 8;
  if (state_bytes(is_matrix, dims, n, pl->es, &pl->bytes) != PW_OK) { delete pl; return PW_ERR_INVALID; }
  for (int i = 0; i < n; ++i) pl->dims[i] = dims[i];
  auto fail = [&](int code) { pw_pipeline_destroy(pl); return code; };
  for (cudaStream_t* s : {&pl->s_h2d, &pl->s_comp, &pl->s_d2h, &pl->s_ops})
    if (cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) != cudaSuccess) return fail(PW_ERR_CUDA);
  for (int i = 0; i < nbuf; ++i) {
    void* p = nullptr;
    if (cudaMalloc(&p, pl->bytes) != cudaSuccess) {
      (void)cudaGetLastError();
      break;  // run with what fits (>= 2)
    }
    pl->owned.push_back(p);
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return fail(PW_ERR_CUDA);
    cudaEventRecord(ev, pl->s_comp);
    pl->pool.push_back({p, ev});
  }
  if (pl->pool.size() < 2) return fail(PW_ERR_CUDA);
  pl->slots.resize(pl->pool.size());
  for (auto& s : pl->slots) {
    if (cudaEventCreateWithFlags(&s.ops_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.done_ev, cudaEventDisableTiming) != cudaSuccess)
      return fail(PW_ERR_CUDA);
  }
  *out = pl;
  return PW_OK;
}

extern "C" int pw_pipeline_buffers(const pw_pipeline* pl) { return pl ? (int)pl->owned.size() : 0; }

extern "C" int pw_pipeline_wait(pw_pipeline* pl, int64_t ticket) {
  if (!pl || ticket < 0 || ticket >= pl->next_ticket) return PW_ERR_INVALID;
  auto& slot = pl->slots[ticket % (int64_t)pl->slots.size()];
  if (slot.ticket != ticket) return PW_OK;  // the slot was reused: that job finished long ago
  PW_CUDA(cudaSetDevice(pl->device));
  PW_CUDA(cudaEventSynchronize(slot.done_ev));
  return PW_OK;
}

extern "C" int pw_pipeline_submit(pw_pipeline* pl, const void* h_state, void* h_out,
                                  const pw_host_op* ops, int nops, int64_t* ticket) {
  if (!pl || !h_state || !h_out || nops < 0 || (nops && !ops)) return PW_ERR_INVALID;
  PW_CUDA(cudaSetDevice(pl->device));
  PW_TRY(check_ops(pl->is_matrix, pl->dims, pl->n, ops, nops));
  const int64_t tk = pl->next_ticket;
  auto& slot = pl->slots[tk % (int64_t)pl->slots.size()];
  if (slot.ticket >= 0) PW_CUDA(cudaEventSynchronize(slot.done_ev));  // ring full: oldest job first

  // operators -> this slot's arena (its own stream, so the small copies never queue behind
  // a 35 GB upload)
  std::vector<size_t> off(nops);
  size_t total = 0;
  for (int i = 0; i < nops; ++i) {
    uint64_t t = 1;
    for (int k = 0; k < ops[i].nt; ++k) t *= (uint64_t)pl->dims[ops[i].targets[k]];
    const int K = ops[i].kind == 2 ? ops[i].K : 1;
    off[i] = total;
    total += (((size_t)K * t * t * pl->es) + 255) & ~(size_t)255;
  }
  if (total > slot.cap) {
    if (slot.arena) PW_CUDA(cudaFree(slot.arena));
    slot.arena = nullptr;
    slot.cap = 0;
    PW_CUDA(cudaMalloc(&slot.arena, total));
    slot.cap = total;
  }
  std::vector<void*> dev_ops(nops);
  for (int i = 0; i < nops; ++i) {
    dev_ops[i] = (char*)slot.arena + off[i];
    const size_t ob = (i + 1 < nops ? off[i + 1] : total) - off[i];
    uint64_t t = 1;
    for (int k = 0; k < ops[i].nt; ++k) t *= (uint64_t)pl->dims[ops[i].targets[k]];
    const size_t exact = (size_t)(ops[i].kind == 2 ? ops[i].K : 1) * t * t * pl->es;
    (void)ob;
    PW_CUDA(cudaMemcpyAsync(dev_ops[i], ops[i].op, exact, cudaMemcpyHostToDevice, pl->s_ops));
  }
  PW_CUDA(cudaEventRecord(slot.ops_ev, pl->s_ops));

  if (pl->pool.size() < 2) return PW_ERR_INVALID;
  pw_pipeline::Buf X = pl->pool.front(); pl->pool.pop_front();
  pw_pipeline::Buf Z = pl->pool.front(); pl->pool.pop_front();
  // Whatever happens below, both buffers go back to the pool: on an early (error) return their
  // "free" events are recorded behind everything already queued on the three streams, so a
  // later job can never touch a buffer that is still being copied.
  struct PoolGuard {
    pw_pipeline* pl; pw_pipeline::Buf a, b; bool armed = true;
    ~PoolGuard() {
      if (!armed) return;
      for (pw_pipeline::Buf* x : {&a, &b}) {
        cudaStreamWaitEvent(pl->s_d2h, x->free_ev, 0);
        cudaEvent_t tmp;
        if (cudaEventCreateWithFlags(&tmp, cudaEventDisableTiming) == cudaSuccess) {
          for (cudaStream_t st : {pl->s_h2d, pl->s_comp}) {
            cudaEventRecord(tmp, st);
            cudaStreamWaitEvent(pl->s_d2h, tmp, 0);
          }
          cudaEventDestroy(tmp);
        }
        cudaEventRecord(x->free_ev, pl->s_d2h);
        pl->pool.push_back(*x);
      }
      (void)cudaGetLastError();
    }
  } guard{pl, X, Z};

  // upload
  PW_CUDA(cudaStreamWaitEvent(pl->s_h2d, X.free_ev, 0));
  PW_CUDA(cudaMemcpyAsync(X.p, h_state, pl->bytes, cudaMemcpyHostToDevice, pl->s_h2d));
  PW_CUDA(cudaEventRecord(X.free_ev, pl->s_h2d));  // reused as "upload done"
  // compute
  PW_CUDA(cudaStreamWaitEvent(pl->s_comp, X.free_ev, 0));
  PW_CUDA(cudaStreamWaitEvent(pl->s_comp, Z.free_ev, 0));
  PW_CUDA(cudaStreamWaitEvent(pl->s_comp, slot.ops_ev, 0));
  void* cur = X.p;
  void* other = Z.p;
  const int fuse = pw::options().fuse_pairs.load(std::memory_order_relaxed);
  const int rc = run_ops(&cur, &other, pl->is_matrix, pl->dims, pl->n, ops, dev_ops.data(), nops,
                         pl->dtype, fuse, pl->s_comp);
  pw_pipeline::Buf R = (cur == X.p) ? X : Z;   // holds the result
  pw_pipeline::Buf O = (cur == X.p) ? Z : X;   // free as soon as the kernels are done
  if (rc != PW_OK) return rc;  // guard returns X and Z
  PW_CUDA(cudaEventRecord(O.free_ev, pl->s_comp));
  // download
  PW_CUDA(cudaStreamWaitEvent(pl->s_d2h, O.free_ev, 0));
  PW_CUDA(cudaMemcpyAsync(h_out, R.p, pl->bytes, cudaMemcpyDeviceToHost, pl->s_d2h));
  PW_CUDA(cudaEventRecord(R.free_ev, pl->s_d2h));
  PW_CUDA(cudaEventRecord(slot.done_ev, pl->s_d2h));
  guard.armed = false;
  pl->pool.push_back(O);
  pl->pool.push_back(R);
  slot.ticket = tk;
  pl->next_ticket = tk + 1;
  if (ticket) *ticket = tk;
  return PW_OK;
}
