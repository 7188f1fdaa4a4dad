// C-ABI entry points for the fused element-wise map and fused map+reduce
// (tile_mapper, spartan/expr/operator/map.py:48-88; _reduce_mapper, reduce.py:21-70).
// Validates the program on the host, then dispatches on the compute dtype.
#include "sp_common.h"
#include <string.h>

namespace sp {
int launch_map_f32(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], cudaStream_t);
int launch_map_f64(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], cudaStream_t);
int launch_map_i64(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], cudaStream_t);
int launch_reduce_f32(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], int, int, void*,
                      int64_t, cudaStream_t);
int launch_reduce_f64(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], int, int, void*,
                      int64_t, cudaStream_t);
int launch_reduce_i64(const sp_program*, int, const sp_operand*, const sp_operand*, const int64_t[3], int, int, void*,
                      int64_t, cudaStream_t);
int64_t reduce_scratch_bytes_f32(const int64_t[3]);
int64_t reduce_scratch_bytes_f64(const int64_t[3]);
int64_t reduce_scratch_bytes_i64(const int64_t[3]);

static bool is_binary(int op) { return op >= SP_OP_ADD && op <= SP_OP_FLOORDIV; }
static bool is_unary(int op) { return (op >= SP_OP_NEG && op <= SP_OP_ISZERO) || (op >= SP_OP_CAST_F32 && op <= SP_OP_CAST_U8); }

// Simulates the stack; the program must leave exactly one value and never exceed the register stack.
static int validate_program(const sp_program* p, int n_in) {
  SP_REQUIRE(p != nullptr, SP_ERR_INVALID, "null program");
  SP_REQUIRE(p->n_ops >= 1 && p->n_ops <= SP_MAX_PROGRAM, SP_ERR_INVALID, "program length %d out of range", p->n_ops);
  SP_REQUIRE(p->compute_dtype == SP_F32 || p->compute_dtype == SP_F64 || p->compute_dtype == SP_I64,
             SP_ERR_UNSUPPORTED, "compute dtype %d not supported", p->compute_dtype);
  SP_REQUIRE(n_in >= 0 && n_in <= SP_MAX_OPERANDS, SP_ERR_UNSUPPORTED, "%d operands exceed the limit of %d", n_in,
             SP_MAX_OPERANDS);
  int sp_ = 0;
  for (int i = 0; i < p->n_ops; ++i) {
    const int op = p->op[i];
    if (op == SP_OP_IN) {
      SP_REQUIRE(p->arg[i] < n_in, SP_ERR_INVALID, "op %d reads operand %d of %d", i, p->arg[i], n_in);
      ++sp_;
    } else if (op == SP_OP_INDEX) {
      ++sp_;
    } else if (op == SP_OP_CONST) {
      SP_REQUIRE(p->arg[i] < SP_MAX_CONSTS, SP_ERR_INVALID, "op %d reads constant %d", i, p->arg[i]);
      ++sp_;
    } else if (is_binary(op)) {
      SP_REQUIRE(sp_ >= 2, SP_ERR_INVALID, "stack underflow at op %d", i);
      --sp_;
    } else if (is_unary(op)) {
      SP_REQUIRE(sp_ >= 1, SP_ERR_INVALID, "stack underflow at op %d", i);
    } else {
      set_error("unknown opcode %d at %d", op, i);
      return SP_ERR_INVALID;
    }
    SP_REQUIRE(sp_ <= 4, SP_ERR_UNSUPPORTED, "expression needs a stack deeper than 4 at op %d", i);
  }
  SP_REQUIRE(sp_ == 1, SP_ERR_INVALID, "program leaves %d values on the stack", sp_);
  return SP_OK;
}

static int check_operands(int n_in, const sp_operand* in, const sp_operand* out, const int64_t dims[3]) {
  SP_REQUIRE(dims[0] >= 0 && dims[1] >= 0 && dims[2] >= 0, SP_ERR_INVALID, "negative dims");
  for (int i = 0; i < n_in; ++i) {
    SP_REQUIRE(in[i].ptr != nullptr, SP_ERR_INVALID, "operand %d is null", i);
    SP_REQUIRE(dtype_size(in[i].dtype) != 0, SP_ERR_INVALID, "operand %d has bad dtype %d", i, in[i].dtype);
  }
  SP_REQUIRE(out != nullptr && out->ptr != nullptr && dtype_size(out->dtype) != 0, SP_ERR_INVALID, "bad output operand");
  return SP_OK;
}

}  // namespace sp

using namespace sp;

extern "C" int sp_map(const sp_program* prog, int n_in, const sp_operand* in, const sp_operand* out,
                      const int64_t dims[3], void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = validate_program(prog, n_in);
  if (rc) return rc;
  rc = check_operands(n_in, in, out, dims);
  if (rc) return rc;
  switch (prog->compute_dtype) {
    case SP_F32: return launch_map_f32(prog, n_in, in, out, dims, stream);
    case SP_F64: return launch_map_f64(prog, n_in, in, out, dims, stream);
    default: return launch_map_i64(prog, n_in, in, out, dims, stream);
  }
}

extern "C" int64_t sp_map_reduce_scratch_bytes(const int64_t dims[3], int compute_dtype) {
  switch (compute_dtype) {
    case SP_F32: return reduce_scratch_bytes_f32(dims);
    case SP_F64: return reduce_scratch_bytes_f64(dims);
    case SP_I64: return reduce_scratch_bytes_i64(dims);
    default: return -1;
  }
}

extern "C" int sp_map_reduce(const sp_program* prog_in, int n_in, const sp_operand* in, const sp_operand* out,
                             const int64_t dims[3], int reduce_op, int accumulate, void* scratch,
                             int64_t scratch_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SP_REQUIRE(prog_in != nullptr, SP_ERR_INVALID, "null program");
  sp_program prog = *prog_in;
  // all / any: fold (x != 0) with min / max  (logic.py:25-46)
  if (reduce_op == SP_RED_ALL || reduce_op == SP_RED_ANY) {
    SP_REQUIRE(prog.n_ops < SP_MAX_PROGRAM, SP_ERR_UNSUPPORTED, "program too long for all/any");
    prog.op[prog.n_ops] = SP_OP_NONZERO;
    prog.arg[prog.n_ops] = 0;
    prog.n_ops++;
    reduce_op = (reduce_op == SP_RED_ALL) ? SP_RED_MIN : SP_RED_MAX;
  }
  SP_REQUIRE(reduce_op >= SP_RED_SUM && reduce_op <= SP_RED_PROD, SP_ERR_INVALID, "bad reduce op %d", reduce_op);
  int rc = validate_program(&prog, n_in);
  if (rc) return rc;
  rc = check_operands(n_in, in, out, dims);
  if (rc) return rc;
  SP_REQUIRE(dims[1] >= 1, SP_ERR_INVALID, "reduction over an empty axis");
  switch (prog.compute_dtype) {
    case SP_F32: return launch_reduce_f32(&prog, n_in, in, out, dims, reduce_op, accumulate, scratch, scratch_bytes, stream);
    case SP_F64: return launch_reduce_f64(&prog, n_in, in, out, dims, reduce_op, accumulate, scratch, scratch_bytes, stream);
    default: return launch_reduce_i64(&prog, n_in, in, out, dims, reduce_op, accumulate, scratch, scratch_bytes, stream);
  }
}
