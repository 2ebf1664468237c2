// Dense one-sided apply on the FP64 tensor cores (see pw_dense.cu).
#pragma once
#include "pw_internal.cuh"

namespace pw {

// out = (O on plan.tm) in for a dense complex128 operator with 8 < t <= 64 whose fibers are
// adjacent in memory (untouched innermost run).  `skip`: optional device flag, the kernel
// exits when it is set.  PW_ERR_UNSUPPORTED when the geometry / dtype does not fit.
int launch_dense_mma(const void* in, void* out, const void* op, const Plan& plan, int dtype,
                     bool conj_op, cudaStream_t stream, const int* skip);

}  // namespace pw
