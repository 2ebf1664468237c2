// Version / status / host-side key management entry points.
#include "pw_internal.cuh"
#include "pw_threefry.h"

#include <cstdlib>
#include <cstring>

extern "C" int pw_version(void) { return PW_B200_VERSION; }

extern "C" const char* pw_status_string(int status) {
  switch (status) {
    case PW_OK: return "ok";
    case PW_ERR_INVALID: return "invalid argument (dims/targets/pointer)";
    case PW_ERR_UNSUPPORTED: return "shape not supported by the kernels";
    case PW_ERR_CUDA: return "CUDA runtime error (see pw_last_cuda_error)";
    case PW_ERR_WORKSPACE: return "stream-ordered workspace allocation failed";
    default: return "unknown status";
  }
}

extern "C" int pw_last_cuda_error(void) { return pw::g_last_cuda_error; }

extern "C" uint64_t pw_launch_count(void) {
  return pw::g_launches.load(std::memory_order_relaxed);
}

// jax.random.split with the partitionable threefry: row i = threefry(key, (hi(i), lo(i))).
extern "C" int pw_rng_split(const uint32_t key[2], int num, uint32_t* out) {
  if (!key || !out || num < 0) return PW_ERR_INVALID;
  for (int i = 0; i < num; ++i)
    pw::threefry2x32(key[0], key[1], 0u, (uint32_t)i, &out[2 * i], &out[2 * i + 1]);
  return PW_OK;
}

extern "C" void pw_path_counters(uint64_t out[5]) {
  for (int i = 0; i < 5; ++i) out[i] = pw::g_paths[i].load(std::memory_order_relaxed);
}

namespace pw {
Options& options() {
  static Options* o = [] {
    Options* p = new Options();
    auto env = [](const char* n, std::atomic<int>& dst) {
      const char* v = std::getenv(n);
      if (v) dst.store(std::atoi(v));
    };
    env("PW_VEC_PATH", p->vec_path);
    env("PW_MAT_MODE", p->mat_mode);
    env("PW_TILE_ELEMS", p->tile_elems);
    env("PW_CTAS_PER_SM", p->ctas_per_sm);
    env("PW_BLOCKS_MIN_ELEMS", p->blocks_min_elems);
    env("PW_BLOCKS_MIN_ELEMS_SMALL_T", p->blocks_min_elems_small_t);
    env("PW_NSTAGE", p->nstage);
    env("PW_SEG_KERNEL", p->seg_kernel);
    env("PW_MAT_PAIR_MAX_D", p->mat_pair_max_d);
    env("PW_FUSE_PAIRS", p->fuse_pairs);
    env("PW_CONTIG_KERNEL", p->contig_kernel);
    env("PW_CONTIG_MMA", p->contig_mma);
    env("PW_DIAG_KERNEL", p->diag_kernel);
    env("PW_SMALL_R", p->small_r);
    env("PW_MAT_SUPER_SINGLE", p->mat_super_single);
    env("PW_TWO_SIDED", p->two_sided);
    env("PW_TS_CTAS", p->ts_ctas);
    env("PW_TS_MMA", p->ts_mma);
    env("PW_TS_MMA_C64", p->ts_mma_c64);
    env("PW_DRAW_EXACT", p->draw_exact);
    env("PW_FUSE_GROUP", p->fuse_group);
    env("PW_MULTI_F", p->multi_f);
    env("PW_HOST_OVERLAP", p->host_overlap);
    env("PW_C64_PAIRS", p->c64_pairs);
    env("PW_TS_ORDER", p->ts_order);
    env("PW_DENSE_MMA", p->dense_mma);
    env("PW_SEG_FPW", p->seg_fpw);
    env("PW_SEG_MMA", p->seg_mma);
    return p;
  }();
  return *o;
}
}  // namespace pw

extern "C" int pw_set_option(const char* name, int value) {
  if (!name) return PW_ERR_INVALID;
  pw::Options& o = pw::options();
  pw::g_options_epoch.fetch_add(1);   // dispatch may change: cached plans of older epochs are not reused
  if (!std::strcmp(name, "vec_path")) o.vec_path.store(value);
  else if (!std::strcmp(name, "mat_mode")) o.mat_mode.store(value);
  else if (!std::strcmp(name, "tile_elems")) o.tile_elems.store(value);
  else if (!std::strcmp(name, "ctas_per_sm")) o.ctas_per_sm.store(value);
  else if (!std::strcmp(name, "blocks_min_elems")) o.blocks_min_elems.store(value);
  else if (!std::strcmp(name, "nstage")) o.nstage.store(value);
  else if (!std::strcmp(name, "seg_kernel")) o.seg_kernel.store(value);
  else if (!std::strcmp(name, "mat_pair_max_d")) o.mat_pair_max_d.store(value);
  else if (!std::strcmp(name, "fuse_pairs")) o.fuse_pairs.store(value);
  else if (!std::strcmp(name, "contig_kernel")) o.contig_kernel.store(value);
  else if (!std::strcmp(name, "small_r")) o.small_r.store(value);
  else if (!std::strcmp(name, "mat_super_single")) o.mat_super_single.store(value);
  else if (!std::strcmp(name, "blocks_min_elems_small_t")) o.blocks_min_elems_small_t.store(value);
  else if (!std::strcmp(name, "contig_mma")) o.contig_mma.store(value);
  else if (!std::strcmp(name, "diag_kernel")) o.diag_kernel.store(value);
  else if (!std::strcmp(name, "two_sided")) o.two_sided.store(value);
  else if (!std::strcmp(name, "ts_ctas")) o.ts_ctas.store(value);
  else if (!std::strcmp(name, "ts_mma")) o.ts_mma.store(value);
  else if (!std::strcmp(name, "ts_mma_c64")) o.ts_mma_c64.store(value);
  else if (!std::strcmp(name, "draw_exact")) o.draw_exact.store(value);
  else if (!std::strcmp(name, "fuse_group")) o.fuse_group.store(value);
  else if (!std::strcmp(name, "multi_f")) o.multi_f.store(value);
  else if (!std::strcmp(name, "host_overlap")) o.host_overlap.store(value);
  else if (!std::strcmp(name, "c64_pairs")) o.c64_pairs.store(value);
  else if (!std::strcmp(name, "ts_order")) o.ts_order.store(value);
  else if (!std::strcmp(name, "dense_mma")) o.dense_mma.store(value);
  else if (!std::strcmp(name, "seg_fpw")) o.seg_fpw.store(value);
  else if (!std::strcmp(name, "seg_mma")) o.seg_mma.store(value);
  else return PW_ERR_INVALID;
  return PW_OK;
}
