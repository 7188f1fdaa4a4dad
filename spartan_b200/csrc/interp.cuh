// In-register bytecode evaluator for fused LocalExpr trees (the fusable kernel IR of the
// reference: spartan/expr/operator/local.py:58-152, fused by optimize.py:133-227).
//
// One thread evaluates the whole program for V elements.  Inputs are loaded up front (all
// loads in flight before the first instruction executes); the program runs on a V-wide
// accumulator (see "the interpreter" below); dispatch cost is amortised over V elements.
#pragma once
#include "sp_common.h"
#ifdef __CUDACC_RTC__
#define CUDART_INF_F __int_as_float(0x7f800000)
#define CUDART_INF __longlong_as_double(0x7ff0000000000000LL)
#else
#include <math_constants.h>
#endif

namespace sp {

template <typename T>
struct DevProgram {
  int32_t n_ops;
  uint8_t op[SP_MAX_PROGRAM];    // accumulator-machine code produced by lower_program()
  uint8_t src[SP_MAX_PROGRAM];
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
// The C ABI carries a postfix (stack) program.  A register-resident *stack* machine turned out to be
// slow on B200: every instruction is a switch case that rewrites part of the V-wide stack, so the
// loop-carried state is the whole stack and ptxas copies ~40 registers per instruction (measured: ~100
// warp-instructions per op, profiles/).  The library therefore lowers the postfix program once per launch
// (lower_program, host) to an ACCUMULATOR machine:
//     acc = x                      A_LOAD        x = in[i] | constant | tmp[k] | element position
//     acc = f(acc, x) / f(x, acc)  binary ops (R* = reversed operands for the non-commutative ones)
//     acc = f(acc)                 unary ops / casts
//     tmp[k] = acc                 A_SAVE        only when both operands of a binary node are sub-trees
// The only loop-carried registers are acc[V]; operands are selected per instruction; tmp lives in local
// memory (L1) and is touched only by A_SAVE / src TMP.  Expression `x*c+y` is 3 instructions.
enum AOp : int {
  A_LOAD = 0, A_SAVE,
  A_ADD, A_SUB, A_MUL, A_DIV, A_MOD, A_POW, A_MAX, A_MIN, A_EQ, A_NE, A_LT, A_LE, A_GT, A_GE, A_AND, A_OR, A_XOR,
  A_FMOD, A_FLOORDIV,
  A_RSUB, A_RDIV, A_RMOD, A_RPOW, A_RFMOD, A_RFLOORDIV,
  A_NEG, A_ABS, A_SQRT, A_EXP, A_LOG, A_SQUARE, A_RECIP, A_NOT, A_NONZERO, A_ISZERO, A_CAST_F32, A_CAST_I64,
  A_CAST_I32, A_CAST_BOOL, A_CAST_U8,
  A_COUNT
};
enum ASrc : int { S_IN0 = 0, S_CONST = 8, S_TMP = 9, S_INDEX = 10, S_NONE = 11 };
constexpr int kMaxTmp = 4;

#ifndef __CUDACC_RTC__   // host-side lowering
struct PostfixNode { int op, arg, l, r; };

inline int aop_binary(int sp_op, bool reversed) {
  switch (sp_op) {
    case SP_OP_ADD: return A_ADD;   case SP_OP_MUL: return A_MUL;   case SP_OP_MAX: return A_MAX;
    case SP_OP_MIN: return A_MIN;   case SP_OP_EQ: return A_EQ;     case SP_OP_NE: return A_NE;
    case SP_OP_AND: return A_AND;   case SP_OP_OR: return A_OR;     case SP_OP_XOR: return A_XOR;
    case SP_OP_SUB: return reversed ? A_RSUB : A_SUB;
    case SP_OP_DIV: return reversed ? A_RDIV : A_DIV;
    case SP_OP_MOD: return reversed ? A_RMOD : A_MOD;
    case SP_OP_POW: return reversed ? A_RPOW : A_POW;
    case SP_OP_FMOD: return reversed ? A_RFMOD : A_FMOD;
    case SP_OP_FLOORDIV: return reversed ? A_RFLOORDIV : A_FLOORDIV;
    case SP_OP_LT: return reversed ? A_GT : A_LT;    // x < acc  <=>  acc > x
    case SP_OP_LE: return reversed ? A_GE : A_LE;
    case SP_OP_GT: return reversed ? A_LT : A_GT;
    case SP_OP_GE: return reversed ? A_LE : A_GE;
    default: return -1;
  }
}
inline int aop_unary(int sp_op) {
  switch (sp_op) {
    case SP_OP_NEG: return A_NEG;       case SP_OP_ABS: return A_ABS;         case SP_OP_SQRT: return A_SQRT;
    case SP_OP_EXP: return A_EXP;       case SP_OP_LOG: return A_LOG;         case SP_OP_SQUARE: return A_SQUARE;
    case SP_OP_RECIP: return A_RECIP;   case SP_OP_NOT: return A_NOT;         case SP_OP_NONZERO: return A_NONZERO;
    case SP_OP_ISZERO: return A_ISZERO; case SP_OP_CAST_F32: return A_CAST_F32;
    case SP_OP_CAST_I64: return A_CAST_I64; case SP_OP_CAST_I32: return A_CAST_I32;
    case SP_OP_CAST_BOOL: return A_CAST_BOOL; case SP_OP_CAST_U8: return A_CAST_U8;
    default: return -1;
  }
}

template <typename T>
struct Lowering {
  DevProgram<T>* out;
  const PostfixNode* nodes;
  bool ok = true;
  int tmp_used = 0;
  void put(int op, int src, int arg) {
    if (out->n_ops >= SP_MAX_PROGRAM) { ok = false; return; }
    out->op[out->n_ops] = static_cast<uint8_t>(op);
    out->src[out->n_ops] = static_cast<uint8_t>(src);
    out->arg[out->n_ops] = static_cast<uint8_t>(arg);
    out->n_ops++;
  }
  bool leaf(int n) const { return nodes[n].op == SP_OP_IN || nodes[n].op == SP_OP_CONST || nodes[n].op == SP_OP_INDEX; }
  void src_of(int n, int* src, int* arg) {
    const PostfixNode& nd = nodes[n];
    if (nd.op == SP_OP_IN) { *src = S_IN0 + nd.arg; *arg = 0; }
    else if (nd.op == SP_OP_CONST) { *src = S_CONST; *arg = nd.arg; }
    else { *src = S_INDEX; *arg = 0; out->uses_index = 1; }
  }
  void emit(int n) {
    if (!ok) return;
    const PostfixNode& nd = nodes[n];
    int src, arg;
    if (leaf(n)) { src_of(n, &src, &arg); put(A_LOAD, src, arg); return; }
    if (nd.r < 0) {                       // unary
      emit(nd.l);
      const int a = aop_unary(nd.op);
      if (a < 0) { ok = false; return; }
      put(a, S_NONE, 0);
      return;
    }
    if (leaf(nd.r)) {                     // acc = f(L, leaf)
      emit(nd.l);
      src_of(nd.r, &src, &arg);
      put(aop_binary(nd.op, false), src, arg);
    } else if (leaf(nd.l)) {              // acc = f(leaf, R)
      emit(nd.r);
      src_of(nd.l, &src, &arg);
      put(aop_binary(nd.op, true), src, arg);
    } else {                              // both sub-trees: park R in a temporary
      emit(nd.r);
      if (tmp_used >= kMaxTmp) { ok = false; return; }
      const int k = tmp_used++;
      put(A_SAVE, S_NONE, k);
      emit(nd.l);
      put(aop_binary(nd.op, false), S_TMP, k);
      tmp_used--;
    }
  }
};

// Host: validated postfix program -> accumulator code.  Returns false if it cannot be lowered.
template <typename T>
inline bool lower_program(const sp_program* prog, DevProgram<T>* out) {
  PostfixNode nodes[SP_MAX_PROGRAM];
  int stack[SP_MAX_PROGRAM];
  int sp = 0;
  for (int i = 0; i < prog->n_ops; ++i) {
    const int op = prog->op[i];
    nodes[i] = PostfixNode{op, prog->arg[i], -1, -1};
    if (op == SP_OP_IN || op == SP_OP_CONST || op == SP_OP_INDEX) {
      stack[sp++] = i;
    } else if (aop_unary(op) >= 0) {
      if (sp < 1) return false;
      nodes[i].l = stack[sp - 1];
      stack[sp - 1] = i;
    } else if (aop_binary(op, false) >= 0) {
      if (sp < 2) return false;
      nodes[i].l = stack[sp - 2];
      nodes[i].r = stack[sp - 1];
      sp -= 1;
      stack[sp - 1] = i;
    } else {
      return false;
    }
  }
  if (sp != 1) return false;
  out->n_ops = 0;
  Lowering<T> lw{out, nodes};
  lw.emit(stack[0]);
  return lw.ok;
}

#endif  // !__CUDACC_RTC__

#define SP_B(x) ((x) ? T(1) : T(0))

// acc[v] = EXPR(a = acc[v], b = operand[v]) with the operand read straight from its home (no staging copy):
// one case per operand source, so every index is a compile-time constant.
#define SP_APPLY(BEXPR, EXPR)                                        \
  {                                                                  \
    _Pragma("unroll") for (int v = 0; v < V; ++v) {                  \
      const T a = acc[v];                                            \
      const T b = (BEXPR);                                           \
      acc[v] = (EXPR);                                               \
    }                                                                \
    break;                                                           \
  }
#define SP_IN(I) in[(I) < NI ? (I) : 0][v]
#define SP_BINARY(AOP, EXPR)                                         \
  case AOP: {                                                        \
    switch (src) {                                                   \
      case 0: SP_APPLY(SP_IN(0), EXPR)                               \
      case 1: SP_APPLY(SP_IN(1), EXPR)                               \
      case 2: SP_APPLY(SP_IN(2), EXPR)                               \
      case 3: SP_APPLY(SP_IN(3), EXPR)                               \
      case 4: SP_APPLY(SP_IN(4), EXPR)                               \
      case 5: SP_APPLY(SP_IN(5), EXPR)                               \
      case 6: SP_APPLY(SP_IN(6), EXPR)                               \
      case 7: SP_APPLY(SP_IN(7), EXPR)                               \
      case S_CONST: SP_APPLY(cst, EXPR)                              \
      case S_TMP: SP_APPLY(tmp[arg][v], EXPR)                        \
      default: SP_APPLY(idx[v], EXPR) /* S_INDEX */                  \
    }                                                                \
    break;                                                           \
  }
// Long-bodied operators (pow, mod, integer division ...) stage the operand once so their code exists once per
// operator instead of once per source.
#define SP_HEAVY(AOP, EXPR)                                          \
  case AOP: {                                                        \
    T x[V];                                                          \
    switch (src) {                                                   \
      case 0: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(0); break; \
      case 1: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(1); break; \
      case 2: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(2); break; \
      case 3: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(3); break; \
      case 4: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(4); break; \
      case 5: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(5); break; \
      case 6: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(6); break; \
      case 7: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = SP_IN(7); break; \
      case S_CONST: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = cst; break; \
      case S_TMP: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = tmp[arg][v]; break; \
      default: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = idx[v]; break; \
    }                                                                \
    SP_APPLY(x[v], EXPR)                                             \
  }
#define SP_UNARY(AOP, EXPR)                                          \
  case AOP: SP_APPLY(T(0), EXPR)

// One instruction.  With compile-time-constant arguments (static programs below) both switches fold away.
template <typename T, int V, int NI>
__device__ __forceinline__ void exec_op(const int op, const int src, const int arg, const DevProgram<T>& prog,
                                        const T (&in)[NI][V], const T (&idx)[V], T (&acc)[V], T (&tmp)[kMaxTmp][V]) {
  const T cst = prog.consts[src == S_CONST ? arg : 0];
  switch (op) {
    SP_BINARY(A_LOAD, b)
    case A_SAVE:
#pragma unroll
      for (int v = 0; v < V; ++v) tmp[arg][v] = acc[v];
      break;
    SP_BINARY(A_ADD, a + b)
    SP_BINARY(A_SUB, a - b)
    SP_BINARY(A_MUL, a * b)
    SP_HEAVY(A_DIV, op_div(a, b))
    SP_HEAVY(A_MOD, op_mod(a, b))
    SP_HEAVY(A_POW, op_pow(a, b))
    SP_BINARY(A_MAX, op_max(a, b))
    SP_BINARY(A_MIN, op_min(a, b))
    SP_BINARY(A_EQ, SP_B(a == b))
    SP_BINARY(A_NE, SP_B(a != b))
    SP_BINARY(A_LT, SP_B(a < b))
    SP_BINARY(A_LE, SP_B(a <= b))
    SP_BINARY(A_GT, SP_B(a > b))
    SP_BINARY(A_GE, SP_B(a >= b))
    SP_BINARY(A_AND, SP_B((a != T(0)) && (b != T(0))))
    SP_BINARY(A_OR, SP_B((a != T(0)) || (b != T(0))))
    SP_BINARY(A_XOR, SP_B((a != T(0)) != (b != T(0))))
    SP_HEAVY(A_FMOD, op_fmod(a, b))
    SP_HEAVY(A_FLOORDIV, op_floordiv(a, b))
    SP_BINARY(A_RSUB, b - a)
    SP_HEAVY(A_RDIV, op_div(b, a))
    SP_HEAVY(A_RMOD, op_mod(b, a))
    SP_HEAVY(A_RPOW, op_pow(b, a))
    SP_HEAVY(A_RFMOD, op_fmod(b, a))
    SP_HEAVY(A_RFLOORDIV, op_floordiv(b, a))
    SP_UNARY(A_NEG, -a)
    SP_UNARY(A_ABS, op_abs(a))
    SP_UNARY(A_SQRT, op_sqrt(a))
    SP_UNARY(A_EXP, op_exp(a))
    SP_UNARY(A_LOG, op_log(a))
    SP_UNARY(A_SQUARE, a * a)
    SP_UNARY(A_RECIP, op_recip(a))
    SP_UNARY(A_NOT, SP_B(a == T(0)))
    SP_UNARY(A_NONZERO, SP_B(a != T(0)))
    SP_UNARY(A_ISZERO, SP_B(a == T(0)))
    SP_UNARY(A_CAST_F32, cast_f32(a))
    SP_UNARY(A_CAST_I64, cast_i64(a))
    SP_UNARY(A_CAST_I32, cast_i32(a))
    SP_UNARY(A_CAST_BOOL, SP_B(a != T(0)))
    SP_UNARY(A_CAST_U8, cast_u8(a))
    default: break;
  }
}

// Evaluates `prog` on the V-wide inputs; result left in out[V].  General path: two small dispatches per instruction.
template <typename T, int V, int NI>
__device__ __forceinline__ void run_program(const DevProgram<T>& prog, const T (&in)[NI][V], const T (&idx)[V],
                                            T (&out)[V]) {
  T tmp[kMaxTmp][V];
  const int n = prog.n_ops;
  for (int pc = 0; pc < n; ++pc) exec_op<T, V, NI>(prog.op[pc], prog.src[pc], prog.arg[pc], prog, in, idx, out, tmp);
}

// ---------------------------------------------------------------------------- static programs
// The hottest fused shapes are also instantiated at compile time: the instruction list is a template
// parameter pack, exec_op is inlined with constant operands, and the kernel body is straight-line code
// (no dispatch at all).  The host matches the lowered program against this catalogue and otherwise uses
// the interpreter above.  PK packs (accumulator op, operand source, argument).
#define SP_PK(AOP, SRC, ARG) (((AOP) << 16) | ((SRC) << 8) | (ARG))

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
    T tmp[kMaxTmp][V];
    (exec_op<T, V, NI>(PKS >> 16, (PKS >> 8) & 0xff, PKS & 0xff, prog, in, idx, out, tmp), ...);
  }
  static bool matches(const uint8_t* op, const uint8_t* src, const uint8_t* arg, int n) {
    const int pk[] = {PKS...};
    if (n != kLen) return false;
    for (int i = 0; i < n; ++i)
      if (((static_cast<int>(op[i]) << 16) | (static_cast<int>(src[i]) << 8) | arg[i]) != pk[i]) return false;
    return true;
  }
};

template <int AOP> using BinaryOf = StaticProgram<SP_PK(A_LOAD, S_IN0, 0), SP_PK(AOP, S_IN0 + 1, 0)>;
template <int AOP> using ScalarOf = StaticProgram<SP_PK(A_LOAD, S_IN0, 0), SP_PK(AOP, S_CONST, 0)>;

// catalogue of statically compiled programs (operands: in0, in1, scalar c0)
using SProg0 = StaticProgram<SP_PK(A_LOAD, S_IN0, 0)>;          // x   (copy, cast, plain reduce)
using SProg1 = BinaryOf<A_ADD>;  using SProg2 = BinaryOf<A_SUB>;  using SProg3 = BinaryOf<A_MUL>;
using SProg4 = BinaryOf<A_DIV>;  using SProg5 = BinaryOf<A_MAX>;  using SProg6 = BinaryOf<A_MIN>;
using SProg7 = ScalarOf<A_ADD>;  using SProg8 = ScalarOf<A_SUB>;  using SProg9 = ScalarOf<A_MUL>;
using SProg10 = ScalarOf<A_DIV>; using SProg11 = ScalarOf<A_MAX>; using SProg12 = ScalarOf<A_MIN>;
// x * c + y -- the fused chain of BASELINE config 3
using SProg13 = StaticProgram<SP_PK(A_LOAD, S_IN0, 0), SP_PK(A_MUL, S_CONST, 0), SP_PK(A_ADD, S_IN0 + 1, 0)>;

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
