// tcgen05 (5th-gen tensor core) grouped GEMM for the typed linears — placeholder until the
// split-bf16 kernel lands; hgt_typed_linear falls back to the fp32 SIMT kernel.
#include "common.cuh"

bool hgt_typed_linear_tc_supported(int64_t lda, int32_t K, int32_t cb_width) {
  (void)lda; (void)K; (void)cb_width;
  return false;
}

int hgt_typed_linear_tc(const float*, int64_t, const float*, const float*, int32_t, int32_t,
                        const hgt_lin_group*, const hgt_lin_group*, int32_t, const hgt_lin_cblock*, float*,
                        cudaStream_t) {
  hgt_set_error("hgt_typed_linear: tensor-core kernel not built");
  return 1;
}
