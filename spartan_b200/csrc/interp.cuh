// In-register bytecode evaluator for fused LocalExpr trees (the fusable kernel IR of the
// reference: spartan/expr/operator/local.py:58-152, fused by optimize.py:133-227).
//
// One thread evaluates the whole program for V consecutive elements.  The operand stack
// lives in registers: every (opcode, stack-depth) pair is its own switch case, so all
// stack indices are compile-time constants and nothing spills to local memory.  The
// dispatch cost (one jump-table branch per op) is amortised over V elements; inputs are
// loaded up front (all loads in flight before the first op executes).
#pragma once
#include "sp_common.h"
#include <math_constants.h>

namespace sp {

constexpr int kMaxDepth = 4;

template <typename T>
struct DevProgram {
  int32_t n_ops;
  uint8_t op[SP_MAX_PROGRAM];
  uint8_t arg[SP_MAX_PROGRAM];
  T consts[SP_MAX_CONSTS];
  long long index_stride[3];
  long long index_base;
  int32_t uses_index;
};

// How an operand is addressed along the vectorised axis.
enum OperandKind : int32_t {
  kVec = 0,      // same dtype as the compute type, unit stride, 16B-aligned vectors
  kSplat = 1,    // stride 0 along the vector axis: one load, broadcast
  kGeneric = 2   // any dtype / stride: element loads with conversion
};

struct DevOperand {
  const void* ptr;
  int32_t dtype;
  int32_t kind;
  int64_t stride[3];
};

template <int NI>
struct DevOperands {
  DevOperand in[NI];
  DevOperand out;
  int32_t n_in;
};

// ---------------------------------------------------------------------------- scalar helpers
template <typename T> struct is_fp { static constexpr bool value = false; };
template <> struct is_fp<float> { static constexpr bool value = true; };
template <> struct is_fp<double> { static constexpr bool value = true; };

template <typename T>
__device__ __forceinline__ T load_as(const void* p, int dtype, int64_t idx) {
  switch (dtype) {
    case SP_F32: return static_cast<T>(static_cast<const float*>(p)[idx]);
    case SP_F64: return static_cast<T>(static_cast<const double*>(p)[idx]);
    case SP_I32: return static_cast<T>(static_cast<const int32_t*>(p)[idx]);
    case SP_I64: return static_cast<T>(static_cast<const long long*>(p)[idx]);
    default:     return static_cast<T>(static_cast<const uint8_t*>(p)[idx]);
  }
}

template <typename T>
__device__ __forceinline__ void store_as(void* p, int dtype, int64_t idx, T v) {
  switch (dtype) {
    case SP_F32: static_cast<float*>(p)[idx] = static_cast<float>(v); break;
    case SP_F64: static_cast<double*>(p)[idx] = static_cast<double>(v); break;
    case SP_I32: static_cast<int32_t*>(p)[idx] = static_cast<int32_t>(v); break;
    case SP_I64: static_cast<long long*>(p)[idx] = static_cast<long long>(v); break;
    case SP_BOOL: static_cast<uint8_t*>(p)[idx] = (v != T(0)) ? 1 : 0; break;
    default:     static_cast<uint8_t*>(p)[idx] = static_cast<uint8_t>(v); break;
  }
}

// np.mod: result has the sign of the divisor
__device__ __forceinline__ float op_mod(float a, float b) { float r = fmodf(a, b); return (r != 0.f && ((r < 0.f) != (b < 0.f))) ? r + b : r; }
__device__ __forceinline__ double op_mod(double a, double b) { double r = fmod(a, b); return (r != 0.0 && ((r < 0.0) != (b < 0.0))) ? r + b : r; }
__device__ __forceinline__ long long op_mod(long long a, long long b) {
  if (b == 0) return 0;  // NumPy: integer x % 0 == 0 (with a warning)
  long long r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? r + b : r;
}
__device__ __forceinline__ float op_fmod(float a, float b) { return fmodf(a, b); }
__device__ __forceinline__ double op_fmod(double a, double b) { return fmod(a, b); }
__device__ __forceinline__ long long op_fmod(long long a, long long b) { return b == 0 ? 0 : a % b; }
__device__ __forceinline__ float op_floordiv(float a, float b) { return floorf(a / b); }
__device__ __forceinline__ double op_floordiv(double a, double b) { return floor(a / b); }
__device__ __forceinline__ long long op_floordiv(long long a, long long b) {
  if (b == 0) return 0;
  long long q = a / b;
  return ((a % b != 0) && ((a < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ float op_div(float a, float b) { return a / b; }
__device__ __forceinline__ double op_div(double a, double b) { return a / b; }
__device__ __forceinline__ long long op_div(long long a, long long b) { return op_floordiv(a, b); }
__device__ __forceinline__ float op_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double op_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ long long op_pow(long long a, long long e) {
  if (e < 0) return 0;
  long long r = 1;
  while (e) { if (e & 1) r *= a; a *= a; e >>= 1; }
  return r;
}
// np.maximum / np.minimum propagate NaN
__device__ __forceinline__ float op_max(float a, float b) { return (a > b || a != a) ? a : b; }
__device__ __forceinline__ double op_max(double a, double b) { return (a > b || a != a) ? a : b; }
__device__ __forceinline__ long long op_max(long long a, long long b) { return a > b ? a : b; }
__device__ __forceinline__ float op_min(float a, float b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ double op_min(double a, double b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ long long op_min(long long a, long long b) { return a < b ? a : b; }

__device__ __forceinline__ float op_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double op_abs(double a) { return fabs(a); }
__device__ __forceinline__ long long op_abs(long long a) { return a < 0 ? -a : a; }
__device__ __forceinline__ float op_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double op_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ long long op_sqrt(long long a) { return static_cast<long long>(sqrt(static_cast<double>(a))); }
__device__ __forceinline__ float op_exp(float a) { return expf(a); }
__device__ __forceinline__ double op_exp(double a) { return exp(a); }
__device__ __forceinline__ long long op_exp(long long a) { return static_cast<long long>(exp(static_cast<double>(a))); }
__device__ __forceinline__ float op_log(float a) { return logf(a); }
__device__ __forceinline__ double op_log(double a) { return log(a); }
__device__ __forceinline__ long long op_log(long long a) { return static_cast<long long>(log(static_cast<double>(a))); }
__device__ __forceinline__ float op_recip(float a) { return 1.0f / a; }
__device__ __forceinline__ double op_recip(double a) { return 1.0 / a; }
__device__ __forceinline__ long long op_recip(long long a) { return a == 0 ? 0 : 1 / a; }

template <typename T> __device__ __forceinline__ T cast_f32(T a) { return static_cast<T>(static_cast<float>(a)); }
template <typename T> __device__ __forceinline__ T cast_i64(T a) { return static_cast<T>(static_cast<long long>(a)); }
template <typename T> __device__ __forceinline__ T cast_i32(T a) { return static_cast<T>(static_cast<int32_t>(static_cast<long long>(a))); }
template <typename T> __device__ __forceinline__ T cast_u8(T a) { return static_cast<T>(static_cast<uint8_t>(static_cast<long long>(a))); }

// ---------------------------------------------------------------------------- the interpreter
// The host lowers the postfix program once more before launch (lower_program below): because the
// stack depth at every instruction is known statically, each instruction becomes ONE dense code
//   code = dense_op * 4 + slot        (slot = index of the stack register the op writes)
// so the kernel dispatches with a single jump table (BRX) and keeps no stack pointer at all.
enum DenseOp : int {
  D_IN = 0, D_CONST, D_INDEX,
  D_ADD, D_SUB, D_MUL, D_DIV, D_MOD, D_POW, D_MAX, D_MIN, D_EQ, D_NE, D_LT, D_LE, D_GT, D_GE, D_AND, D_OR, D_XOR,
  D_FMOD, D_FLOORDIV,
  D_NEG, D_ABS, D_SQRT, D_EXP, D_LOG, D_SQUARE, D_RECIP, D_NOT, D_NONZERO, D_ISZERO, D_CAST_F32, D_CAST_I64,
  D_CAST_I32, D_CAST_BOOL, D_CAST_U8,
  D_COUNT
};

inline int dense_of(int sp_op) {
  switch (sp_op) {
    case SP_OP_IN: return D_IN;          case SP_OP_CONST: return D_CONST;   case SP_OP_INDEX: return D_INDEX;
    case SP_OP_ADD: return D_ADD;        case SP_OP_SUB: return D_SUB;       case SP_OP_MUL: return D_MUL;
    case SP_OP_DIV: return D_DIV;        case SP_OP_MOD: return D_MOD;       case SP_OP_POW: return D_POW;
    case SP_OP_MAX: return D_MAX;        case SP_OP_MIN: return D_MIN;       case SP_OP_EQ: return D_EQ;
    case SP_OP_NE: return D_NE;          case SP_OP_LT: return D_LT;         case SP_OP_LE: return D_LE;
    case SP_OP_GT: return D_GT;          case SP_OP_GE: return D_GE;         case SP_OP_AND: return D_AND;
    case SP_OP_OR: return D_OR;          case SP_OP_XOR: return D_XOR;       case SP_OP_FMOD: return D_FMOD;
    case SP_OP_FLOORDIV: return D_FLOORDIV;
    case SP_OP_NEG: return D_NEG;        case SP_OP_ABS: return D_ABS;       case SP_OP_SQRT: return D_SQRT;
    case SP_OP_EXP: return D_EXP;        case SP_OP_LOG: return D_LOG;       case SP_OP_SQUARE: return D_SQUARE;
    case SP_OP_RECIP: return D_RECIP;    case SP_OP_NOT: return D_NOT;       case SP_OP_NONZERO: return D_NONZERO;
    case SP_OP_ISZERO: return D_ISZERO;  case SP_OP_CAST_F32: return D_CAST_F32;
    case SP_OP_CAST_I64: return D_CAST_I64; case SP_OP_CAST_I32: return D_CAST_I32;
    case SP_OP_CAST_BOOL: return D_CAST_BOOL; case SP_OP_CAST_U8: return D_CAST_U8;
    default: return -1;
  }
}

// Host: validated postfix program -> dense codes.  Returns false if the program is malformed.
template <typename T>
inline bool lower_program(const sp_program* prog, DevProgram<T>* out) {
  out->n_ops = prog->n_ops;
  int sp = 0;
  for (int i = 0; i < prog->n_ops; ++i) {
    const int d = dense_of(prog->op[i]);
    if (d < 0) return false;
    int slot;
    if (d <= D_INDEX) { slot = sp; sp += 1; if (d == D_INDEX) out->uses_index = 1; }
    else if (d <= D_FLOORDIV) { slot = sp - 2; sp -= 1; }
    else { slot = sp - 1; }
    if (slot < 0 || slot >= kMaxDepth) return false;
    out->op[i] = static_cast<uint8_t>(d * 4 + slot);
    out->arg[i] = prog->arg[i];
  }
  return sp == 1;
}

#define SP_B(x) ((x) ? T(1) : T(0))

#define SP_BIN_CASE(DOP, S, EXPR)                                    \
  case (DOP) * 4 + (S): {                                            \
    _Pragma("unroll") for (int v = 0; v < V; ++v) {                  \
      const T a = s[(S)][v];                                         \
      const T b = s[(S) + 1][v];                                     \
      s[(S)][v] = (EXPR);                                            \
    }                                                                \
    break;                                                           \
  }
#define SP_BIN(DOP, EXPR) SP_BIN_CASE(DOP, 0, EXPR) SP_BIN_CASE(DOP, 1, EXPR) SP_BIN_CASE(DOP, 2, EXPR)

#define SP_UN_CASE(DOP, S, EXPR)                                     \
  case (DOP) * 4 + (S): {                                            \
    _Pragma("unroll") for (int v = 0; v < V; ++v) {                  \
      const T a = s[(S)][v];                                         \
      s[(S)][v] = (EXPR);                                            \
    }                                                                \
    break;                                                           \
  }
#define SP_UN(DOP, EXPR) SP_UN_CASE(DOP, 0, EXPR) SP_UN_CASE(DOP, 1, EXPR) SP_UN_CASE(DOP, 2, EXPR) SP_UN_CASE(DOP, 3, EXPR)

#define SP_PUSH_CONST_CASE(S)                                        \
  case D_CONST * 4 + (S): {                                          \
    const T c = prog.consts[arg];                                    \
    _Pragma("unroll") for (int v = 0; v < V; ++v) s[(S)][v] = c;     \
    break;                                                           \
  }

#define SP_PUSH_INDEX_CASE(S)                                        \
  case D_INDEX * 4 + (S): {                                          \
    _Pragma("unroll") for (int v = 0; v < V; ++v) s[(S)][v] = idx[v]; \
    break;                                                           \
  }

// push operand: the operand index is data, the destination register is static
#define SP_PUSH_IN_CASE(S)                                           \
  case D_IN * 4 + (S): {                                             \
    _Pragma("unroll") for (int i = 0; i < NI; ++i) {                 \
      if (i == arg) {                                                \
        _Pragma("unroll") for (int v = 0; v < V; ++v) s[(S)][v] = in[i][v]; \
      }                                                              \
    }                                                                \
    break;                                                           \
  }

// One instruction.  With compile-time-constant `code` / `arg` (static programs below) the switch folds
// away and only the selected case's V arithmetic instructions remain.
template <typename T, int V, int NI>
__device__ __forceinline__ void exec_op(const int code, const int arg, const DevProgram<T>& prog, const T (&in)[NI][V],
                                        const T (&idx)[V], T (&s)[kMaxDepth][V]) {
  switch (code) {
      SP_PUSH_INDEX_CASE(0) SP_PUSH_INDEX_CASE(1) SP_PUSH_INDEX_CASE(2) SP_PUSH_INDEX_CASE(3)
      SP_PUSH_IN_CASE(0) SP_PUSH_IN_CASE(1) SP_PUSH_IN_CASE(2) SP_PUSH_IN_CASE(3)
      SP_PUSH_CONST_CASE(0) SP_PUSH_CONST_CASE(1) SP_PUSH_CONST_CASE(2) SP_PUSH_CONST_CASE(3)
      SP_BIN(D_ADD, a + b)
      SP_BIN(D_SUB, a - b)
      SP_BIN(D_MUL, a * b)
      SP_BIN(D_DIV, op_div(a, b))
      SP_BIN(D_MOD, op_mod(a, b))
      SP_BIN(D_POW, op_pow(a, b))
      SP_BIN(D_MAX, op_max(a, b))
      SP_BIN(D_MIN, op_min(a, b))
      SP_BIN(D_EQ, SP_B(a == b))
      SP_BIN(D_NE, SP_B(a != b))
      SP_BIN(D_LT, SP_B(a < b))
      SP_BIN(D_LE, SP_B(a <= b))
      SP_BIN(D_GT, SP_B(a > b))
      SP_BIN(D_GE, SP_B(a >= b))
      SP_BIN(D_AND, SP_B((a != T(0)) && (b != T(0))))
      SP_BIN(D_OR, SP_B((a != T(0)) || (b != T(0))))
      SP_BIN(D_XOR, SP_B((a != T(0)) != (b != T(0))))
      SP_BIN(D_FMOD, op_fmod(a, b))
      SP_BIN(D_FLOORDIV, op_floordiv(a, b))
      SP_UN(D_NEG, -a)
      SP_UN(D_ABS, op_abs(a))
      SP_UN(D_SQRT, op_sqrt(a))
      SP_UN(D_EXP, op_exp(a))
      SP_UN(D_LOG, op_log(a))
      SP_UN(D_SQUARE, a * a)
      SP_UN(D_RECIP, op_recip(a))
      SP_UN(D_NOT, SP_B(a == T(0)))
      SP_UN(D_NONZERO, SP_B(a != T(0)))
      SP_UN(D_ISZERO, SP_B(a == T(0)))
      SP_UN(D_CAST_F32, cast_f32(a))
      SP_UN(D_CAST_I64, cast_i64(a))
      SP_UN(D_CAST_I32, cast_i32(a))
      SP_UN(D_CAST_BOOL, SP_B(a != T(0)))
      SP_UN(D_CAST_U8, cast_u8(a))
    default: break;   // slot 3 of a binary op: unreachable for validated programs
  }
}

// Evaluates `prog` on the V-wide inputs; result left in out[V].  General path: one jump-table dispatch per op.
template <typename T, int V, int NI>
__device__ __forceinline__ void run_program(const DevProgram<T>& prog, const T (&in)[NI][V], const T (&idx)[V],
                                            T (&out)[V]) {
  T s[kMaxDepth][V];
  const int n = prog.n_ops;
  for (int pc = 0; pc < n; ++pc) exec_op<T, V, NI>(prog.op[pc], prog.arg[pc], prog, in, idx, s);
#pragma unroll
  for (int v = 0; v < V; ++v) out[v] = s[0][v];
}

// ---------------------------------------------------------------------------- static programs
// The hottest fused shapes are also instantiated at compile time: the instruction list is a template
// parameter pack, exec_op is inlined with constant operands, and the kernel body is straight-line code
// (no dispatch at all).  The host matches the lowered program against this catalogue and otherwise uses
// the interpreter above.  PK packs (dense op, stack slot, argument).
#define SP_PK(DOP, SLOT, ARG) ((((DOP) * 4 + (SLOT)) << 8) | (ARG))

struct DynamicProgram {
  template <typename T, int V, int NI>
  static __device__ __forceinline__ void run(const DevProgram<T>& prog, const T (&in)[NI][V], const T (&idx)[V],
                                             T (&out)[V]) {
    run_program<T, V, NI>(prog, in, idx, out);
  }
};

template <int... PKS>
struct StaticProgram {
  static constexpr int kLen = sizeof...(PKS);
  template <typename T, int V, int NI>
  static __device__ __forceinline__ void run(const DevProgram<T>& prog, const T (&in)[NI][V], const T (&idx)[V],
                                             T (&out)[V]) {
    T s[kMaxDepth][V];
    (exec_op<T, V, NI>(PKS >> 8, PKS & 0xff, prog, in, idx, s), ...);
#pragma unroll
    for (int v = 0; v < V; ++v) out[v] = s[0][v];
  }
  static bool matches(const uint8_t* op, const uint8_t* arg, int n) {
    const int pk[] = {PKS...};
    if (n != kLen) return false;
    for (int i = 0; i < n; ++i)
      if (((static_cast<int>(op[i]) << 8) | arg[i]) != pk[i]) return false;
    return true;
  }
};

template <int DOP> using BinaryOf = StaticProgram<SP_PK(D_IN, 0, 0), SP_PK(D_IN, 1, 1), SP_PK(DOP, 0, 0)>;
template <int DOP> using ScalarOf = StaticProgram<SP_PK(D_IN, 0, 0), SP_PK(D_CONST, 1, 0), SP_PK(DOP, 0, 0)>;

// catalogue of statically compiled programs (operands: in0, in1, scalar c0)
using SProg0 = StaticProgram<SP_PK(D_IN, 0, 0)>;          // x   (copy, cast, plain reduce)
using SProg1 = BinaryOf<D_ADD>;  using SProg2 = BinaryOf<D_SUB>;  using SProg3 = BinaryOf<D_MUL>;
using SProg4 = BinaryOf<D_DIV>;  using SProg5 = BinaryOf<D_MAX>;  using SProg6 = BinaryOf<D_MIN>;
using SProg7 = ScalarOf<D_ADD>;  using SProg8 = ScalarOf<D_SUB>;  using SProg9 = ScalarOf<D_MUL>;
using SProg10 = ScalarOf<D_DIV>; using SProg11 = ScalarOf<D_MAX>; using SProg12 = ScalarOf<D_MIN>;
// x * c + y -- the fused chain of BASELINE config 3  (x * y + z has three operands: interpreter, 8-operand variant)
using SProg13 = StaticProgram<SP_PK(D_IN, 0, 0), SP_PK(D_CONST, 1, 0), SP_PK(D_MUL, 0, 0), SP_PK(D_IN, 1, 1),
                              SP_PK(D_ADD, 0, 0)>;

#define SP_STATIC_PROGRAMS(X)                                                                              \
  X(0, SProg0) X(1, SProg1) X(2, SProg2) X(3, SProg3) X(4, SProg4) X(5, SProg5) X(6, SProg6) X(7, SProg7)   \
  X(8, SProg8) X(9, SProg9) X(10, SProg10) X(11, SProg11) X(12, SProg12) X(13, SProg13)

constexpr int kNumStaticPrograms = 14;

// Position operand of SP_OP_INDEX for V lanes: lane v sits at coordinate c_vec + v*step along `vec_axis`.
template <typename T, int V>
__device__ __forceinline__ void make_index(const DevProgram<T>& prog, long long i0, long long i1, long long i2,
                                           int vec_axis, T (&idx)[V]) {
  const long long base = prog.index_base + i0 * prog.index_stride[0] + i1 * prog.index_stride[1] + i2 * prog.index_stride[2];
  const long long step = prog.index_stride[vec_axis];
#pragma unroll
  for (int v = 0; v < V; ++v) idx[v] = static_cast<T>(base + v * step);
}

// ---------------------------------------------------------------------------- operand access
template <typename T, int V> struct VecLoad;
template <> struct VecLoad<float, 8> {
  static __device__ __forceinline__ void ld(const float* p, float (&r)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&r)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
};
template <> struct VecLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&r)[4]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&r)[4]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
  }
};
template <typename T> struct VecLoad<T, 4> {   // 8-byte types
  static __device__ __forceinline__ void ld(const T* p, T (&r)[4]) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(p));
    const longlong2 b = __ldg(reinterpret_cast<const longlong2*>(p) + 1);
    long long t[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = *reinterpret_cast<T*>(&t[i]);
  }
  static __device__ __forceinline__ void st(T* p, const T (&r)[4]) {
    long long t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = *reinterpret_cast<const long long*>(&r[i]);
    reinterpret_cast<longlong2*>(p)[0] = make_longlong2(t[0], t[1]);
    reinterpret_cast<longlong2*>(p)[1] = make_longlong2(t[2], t[3]);
  }
};
template <typename T> struct VecLoad<T, 2> {
  static __device__ __forceinline__ void ld(const T* p, T (&r)[2]) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(p));
    long long t[2] = {a.x, a.y};
    r[0] = *reinterpret_cast<T*>(&t[0]);
    r[1] = *reinterpret_cast<T*>(&t[1]);
  }
  static __device__ __forceinline__ void st(T* p, const T (&r)[2]) {
    reinterpret_cast<longlong2*>(p)[0] =
        make_longlong2(*reinterpret_cast<const long long*>(&r[0]), *reinterpret_cast<const long long*>(&r[1]));
  }
};

// Loads V elements of operand `o` starting at element offset `base` (already includes the d0/d1
// terms), stepping `vstride` elements along the vector axis; `valid` = number of in-range lanes.
template <typename T, int V>
__device__ __forceinline__ void load_operand(const DevOperand& o, int64_t base, int64_t vstride, int valid, T (&r)[V]) {
  if (o.kind == kVec && valid == V) {
    VecLoad<T, V>::ld(static_cast<const T*>(o.ptr) + base, r);
  } else if (o.kind == kSplat) {
    const T x = load_as<T>(o.ptr, o.dtype, base);
#pragma unroll
    for (int v = 0; v < V; ++v) r[v] = x;
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) r[v] = (v < valid) ? load_as<T>(o.ptr, o.dtype, base + v * vstride) : T(0);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_operand(const DevOperand& o, int64_t base, int64_t vstride, int valid, const T (&r)[V]) {
  if (o.kind == kVec && valid == V) {
    VecLoad<T, V>::st(static_cast<T*>(const_cast<void*>(o.ptr)) + base, r);
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (v < valid) store_as<T>(const_cast<void*>(o.ptr), o.dtype, base + v * vstride, r[v]);
  }
}

// ---------------------------------------------------------------------------- reductions
template <typename T> __device__ __forceinline__ T red_identity(int op);
template <> __device__ __forceinline__ float red_identity<float>(int op) {
  return op == SP_RED_SUM ? 0.f : op == SP_RED_PROD ? 1.f : op == SP_RED_MIN ? CUDART_INF_F : -CUDART_INF_F;
}
template <> __device__ __forceinline__ double red_identity<double>(int op) {
  return op == SP_RED_SUM ? 0.0 : op == SP_RED_PROD ? 1.0 : op == SP_RED_MIN ? CUDART_INF : -CUDART_INF;
}
template <> __device__ __forceinline__ long long red_identity<long long>(int op) {
  return op == SP_RED_SUM ? 0ll : op == SP_RED_PROD ? 1ll : op == SP_RED_MIN ? 0x7fffffffffffffffll : (-0x7fffffffffffffffll - 1);
}

template <typename T>
__device__ __forceinline__ T red_apply(int op, T a, T b) {
  switch (op) {
    case SP_RED_SUM: return a + b;
    case SP_RED_PROD: return a * b;
    case SP_RED_MIN: return op_min(a, b);
    default: return op_max(a, b);
  }
}

template <typename T>
__device__ __forceinline__ T shfl_down_t(T v, int delta) {
  if constexpr (sizeof(T) == 4) {
    return __shfl_down_sync(0xffffffffu, v, delta);
  } else {
    long long x = *reinterpret_cast<long long*>(&v);
    x = __shfl_down_sync(0xffffffffu, x, delta);
    return *reinterpret_cast<T*>(&x);
  }
}

}  // namespace sp
