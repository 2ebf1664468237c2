// Version / status / host-side key management entry points.
#include "pw_internal.cuh"
#include "pw_threefry.h"

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
