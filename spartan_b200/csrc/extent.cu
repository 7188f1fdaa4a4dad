// Host-only extent algebra and tiling (int64, bit-exact with the reference's Cython module).
// Native counterpart of spartan/array/extent.pyx and of the tiling helpers in
// spartan/array/distarray.py:26-110.  No device code; lives in a .cu only so the whole shim
// is one translation-unit family built by nvcc.
#include "sp_common.h"
#include <math.h>
#include <vector>

namespace sp {

// spartan/util.py:404-408 : int(ceil(float(a) / b)) -- float ceil, exact below 2^53
static inline int64_t divup_float(int64_t a, int64_t b) {
  return static_cast<int64_t>(ceil(static_cast<double>(a) / static_cast<double>(b)));
}

}  // namespace sp

using namespace sp;

// extent.pyx:367-387 + create() validity (:141-153): returns 1 valid, 0 empty ("None")
extern "C" int sp_extent_intersection(int ndim, const int64_t* a_ul, const int64_t* a_lr, const int64_t* b_ul,
                                      const int64_t* b_lr, int64_t* out_ul, int64_t* out_lr) {
  SP_REQUIRE(ndim >= 0 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  for (int i = 0; i < ndim; ++i) {
    if (b_lr[i] < a_ul[i]) return 0;
    if (a_lr[i] < b_ul[i]) return 0;
    out_ul[i] = a_ul[i] >= b_ul[i] ? a_ul[i] : b_ul[i];
    out_lr[i] = a_lr[i] < b_lr[i] ? a_lr[i] : b_lr[i];
  }
  for (int i = 0; i < ndim; ++i)
    if (out_ul[i] >= out_lr[i]) return 0;
  return 1;
}

// extent.pyx:207-219
extern "C" int64_t sp_extent_ravelled_pos(int ndim, const int64_t* idx, const int64_t* array_shape) {
  int64_t rpos = 0, mul = 1;
  for (int i = ndim - 1; i >= 0; --i) {
    rpos += mul * idx[i];
    mul *= array_shape[i];
  }
  return rpos;
}

// extent.pyx:196-205 (Python-2 integer division)
extern "C" int sp_extent_unravelled_pos(int64_t idx, int ndim, const int64_t* array_shape, int64_t* out_idx) {
  SP_REQUIRE(ndim >= 0 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  for (int i = ndim - 1; i >= 0; --i) {
    const int64_t dim = array_shape[i];
    SP_REQUIRE(dim > 0, SP_ERR_INVALID, "zero-sized dimension");
    int64_t q = idx / dim, r = idx % dim;
    if (r < 0) { r += dim; q -= 1; }   // Python floor semantics
    out_idx[i] = r;
    idx = q;
  }
  return SP_OK;
}

// extent.pyx:411-432
extern "C" int sp_extent_drop_axis(int ndim, const int64_t* ul, const int64_t* lr, const int64_t* array_shape,
                                   int axis, int64_t* out_ul, int64_t* out_lr, int64_t* out_shape, int* out_ndim) {
  SP_REQUIRE(ndim >= 0 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  if (axis == SP_AXIS_NONE) {
    *out_ndim = 0;
    return 1;
  }
  if (axis < 0) axis += ndim;
  SP_REQUIRE(axis >= 0 && axis < ndim, SP_ERR_INVALID, "axis out of range");
  int k = 0;
  for (int i = 0; i < ndim; ++i) {
    if (i == axis) continue;
    out_ul[k] = ul[i];
    out_lr[k] = lr[i];
    out_shape[k] = array_shape[i];
    ++k;
  }
  *out_ndim = ndim - 1;
  for (int i = 0; i < k; ++i)
    if (out_ul[i] >= out_lr[i]) return 0;
  return 1;
}

// extent.pyx:121-127
extern "C" int64_t sp_extent_to_global(int ndim, const int64_t* ul, const int64_t* lr, const int64_t* array_shape,
                                       int64_t idx, int axis) {
  if (axis != SP_AXIS_NONE) return idx + ul[axis];
  if (ndim < 0 || ndim > SP_MAX_DIM) return -1;
  int64_t shape[SP_MAX_DIM] = {0}, local[SP_MAX_DIM] = {0}, glob[SP_MAX_DIM] = {0};
  for (int i = 0; i < ndim; ++i) {
    shape[i] = lr[i] - ul[i];
    if (shape[i] == 0) shape[i] = 1;   // extent.pyx:66-72
  }
  sp_extent_unravelled_pos(idx, ndim, shape, local);
  for (int i = 0; i < ndim; ++i) glob[i] = ul[i] + local[i];
  return sp_extent_ravelled_pos(ndim, glob, array_shape);
}

// extent.pyx:501-570, one-dimensional target
extern "C" int sp_extent_change_partition_axis(int ndim, const int64_t* ul, const int64_t* lr,
                                               const int64_t* array_shape, int axis, int64_t* out_ul,
                                               int64_t* out_lr) {
  SP_REQUIRE(ndim >= 1 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  if (axis < 0) axis += ndim;
  if (ndim == 1) {   // :533-539 vectors
    if (axis == 1) {
      out_ul[0] = 0;
      out_lr[0] = array_shape[0];
    } else {
      out_ul[0] = ul[0];
      out_lr[0] = lr[0];
    }
    return out_ul[0] < out_lr[0] ? 1 : 0;
  }
  SP_REQUIRE(axis >= 0 && axis < ndim, SP_ERR_INVALID, "axis out of range");
  int old_axis = -1, n_part = 0;
  for (int i = 0; i < ndim; ++i) {
    int64_t s = lr[i] - ul[i];
    if (s == 0) s = 1;
    if (s != array_shape[i]) {
      ++n_part;
      if (old_axis < 0) old_axis = i;
    }
  }
  if (n_part > 1) {
    set_error("change_partition_axis on a grid-tiled extent is a reference defect (extent.pyx:545-552); "
              "the caller must handle grid tilings explicitly");
    return SP_ERR_UNSUPPORTED;
  }
  for (int i = 0; i < ndim; ++i) {
    out_ul[i] = ul[i];
    out_lr[i] = lr[i];
  }
  if (n_part == 0 || old_axis == axis) return 1;   // :554-555
  out_ul[axis] = divup_float(ul[old_axis] * array_shape[axis], array_shape[old_axis]);
  out_ul[old_axis] = 0;
  out_lr[axis] = divup_float(lr[old_axis] * array_shape[axis], array_shape[old_axis]);
  out_lr[old_axis] = array_shape[old_axis];
  for (int i = 0; i < ndim; ++i)
    if (out_ul[i] >= out_lr[i]) return 0;
  return 1;
}

// distarray.py:26-48
extern "C" int sp_good_tile_shape(int ndim, const int64_t* shape, int64_t num_shards, int64_t* out_tile_shape) {
  SP_REQUIRE(ndim >= 0 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  int64_t tile_size;
  if (num_shards != -1) {
    SP_REQUIRE(num_shards > 0, SP_ERR_INVALID, "num_shards must be positive or -1");
    int64_t prod = 1;
    for (int i = 0; i < ndim; ++i) prod *= shape[i];
    tile_size = prod / num_shards;
  } else {
    tile_size = 100000;   // DEFAULT_TILE_SIZE, distarray.py:20
  }
  for (int i = 0; i < ndim; ++i) out_tile_shape[i] = 1;
  int idx = ndim - 1;
  while (tile_size > 1) {
    SP_REQUIRE(idx >= 0, SP_ERR_INVALID, "good_tile_shape ran out of dimensions");   // IndexError in the reference
    out_tile_shape[idx] = shape[idx] < tile_size ? shape[idx] : tile_size;
    tile_size /= shape[idx];
    --idx;
  }
  return SP_OK;
}

// distarray.py:51-110
extern "C" int64_t sp_compute_extents(int ndim, const int64_t* shape, const int64_t* tile_hint, int64_t num_shards,
                                      int64_t* out_ul, int64_t* out_lr, int64_t* out_worker) {
  SP_REQUIRE(ndim >= 0 && ndim <= SP_MAX_DIM, SP_ERR_INVALID, "ndim %d out of range", ndim);
  if (ndim == 0) {   // :87-88 the single 0-d extent
    if (out_worker) out_worker[0] = 0;
    return 1;
  }
  int64_t hint[SP_MAX_DIM];
  if (tile_hint == nullptr) {
    const int rc = sp_good_tile_shape(ndim, shape, num_shards, hint);
    if (rc) return rc;
  } else {
    for (int i = 0; i < ndim; ++i) {
      SP_REQUIRE(tile_hint[i] > 0, SP_ERR_INVALID, "tile_hint[%d] must be positive", i);
      hint[i] = tile_hint[i];
    }
  }
  int64_t counts[SP_MAX_DIM], total = 1;
  for (int i = 0; i < ndim; ++i) {
    counts[i] = (shape[i] + hint[i] - 1) / hint[i];   // len(range(0, shape, step))
    total *= counts[i];
  }
  if (out_ul == nullptr) return total;
  int64_t pos[SP_MAX_DIM] = {0};
  int64_t idx = 0;
  for (int64_t t = 0; t < total; ++t) {
    if (num_shards != -1) idx = idx % num_shards;
    for (int i = 0; i < ndim; ++i) {
      const int64_t lo = pos[i] * hint[i];
      const int64_t hi = lo + hint[i] < shape[i] ? lo + hint[i] : shape[i];
      out_ul[t * ndim + i] = lo;
      out_lr[t * ndim + i] = hi;
    }
    out_worker[t] = idx;
    ++idx;
    for (int i = ndim - 1; i >= 0; --i) {   // itertools.product order: last dimension fastest
      if (++pos[i] < counts[i]) break;
      pos[i] = 0;
    }
  }
  return total;
}
