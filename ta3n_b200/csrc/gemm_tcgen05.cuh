// gemm_tcgen05.cuh -- tcgen05 (5th-gen tensor core) engine for the segmented grouped GEMM.
// Placeholder until the engine lands: declines every plan so the SIMT engine runs it.
#pragma once
#include "seg_gemm.cuh"

namespace ta3n {
inline int run_gemm_tcgen05(const GemmPlan& plan, cudaStream_t stream, bool* handled) {
  (void)plan;
  (void)stream;
  *handled = false;
  return TA3N_OK;
}
}  // namespace ta3n
