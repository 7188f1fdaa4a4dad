// Instantiates the fused map / map+reduce kernels for compute type `float` (see map_reduce_impl.cuh).
#include "map_reduce_impl.cuh"
#include <algorithm>
#include <string.h>

namespace sp {
int launch_map_f32(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                  const int64_t dims[3], cudaStream_t stream) {
  return launch_map<float>(prog, n_in, in, out, dims, stream);
}
int launch_reduce_f32(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                     const int64_t dims[3], int red_op, int accumulate, void* scratch, int64_t scratch_bytes,
                     cudaStream_t stream) {
  return launch_reduce<float>(prog, n_in, in, out, dims, red_op, accumulate, scratch, scratch_bytes, stream);
}
int64_t reduce_scratch_bytes_f32(const int64_t dims[3]) { return reduce_scratch_bytes<float>(dims); }

// Accumulator-machine code of a postfix program (the lowering is the same for every compute type): what jit.cu
// keys its specialisations on.
int lower_for_jit(const sp_program* prog, uint8_t* op, uint8_t* src, uint8_t* arg, int* n) {
  DevProgram<float> dp;
  memset(&dp, 0, sizeof(dp));
  SP_REQUIRE(prog->n_ops >= 1 && prog->n_ops <= SP_MAX_PROGRAM && lower_program<float>(prog, &dp), SP_ERR_UNSUPPORTED,
             "program cannot be lowered to the accumulator machine");
  memcpy(op, dp.op, dp.n_ops); memcpy(src, dp.src, dp.n_ops); memcpy(arg, dp.arg, dp.n_ops);
  *n = dp.n_ops;
  return SP_OK;
}
}  // namespace sp
