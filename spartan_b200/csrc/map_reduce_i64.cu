// Instantiates the fused map / map+reduce kernels for compute type `long long` (see map_reduce_impl.cuh).
#include "map_reduce_impl.cuh"
#include <algorithm>
#include <string.h>

namespace sp {
int launch_map_i64(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                  const int64_t dims[3], cudaStream_t stream) {
  return launch_map<long long>(prog, n_in, in, out, dims, stream);
}
int launch_reduce_i64(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                     const int64_t dims[3], int red_op, int accumulate, void* scratch, int64_t scratch_bytes,
                     cudaStream_t stream) {
  return launch_reduce<long long>(prog, n_in, in, out, dims, red_op, accumulate, scratch, scratch_bytes, stream);
}
int64_t reduce_scratch_bytes_i64(const int64_t dims[3]) { return reduce_scratch_bytes<long long>(dims); }
}  // namespace sp
