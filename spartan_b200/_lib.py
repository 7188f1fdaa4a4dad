"""ctypes binding of libspartan_b200.so (the C ABI in include/spartan_b200.h).

The library is the only compute backend: if it cannot be loaded the import fails loudly, and every
compute entry point fails with SP_ERR_CUDA when there is no CUDA device -- there is no CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libspartan_b200.so')

if not os.path.exists(LIB_PATH):
  raise ImportError('%s is missing: build it with `make -C spartan_b200/csrc` (or __graft_entry__.build()); '
                    'spartan_b200 has no CPU fallback' % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

# enums (include/spartan_b200.h)
SP_F32, SP_F64, SP_I32, SP_I64, SP_U8, SP_BOOL = range(6)
SP_AXIS_NONE = -1000
SP_FILL_CONST, SP_FILL_IOTA, SP_FILL_RAND, SP_FILL_RANDN = range(4)
SP_RED_SUM, SP_RED_MIN, SP_RED_MAX, SP_RED_PROD, SP_RED_ALL, SP_RED_ANY = range(6)
SP_GEMM_TF32X1, SP_GEMM_TF32X3, SP_GEMM_SIMT, SP_GEMM_BF16X3 = range(4)
SP_GEMM_MAX_SEGMENTS = 8
SP_MAX_PROGRAM, SP_MAX_OPERANDS, SP_MAX_CONSTS, SP_MAX_STACK = 64, 8, 16, 4
SP_GEMM_MAX_TERMS = 24

OP = dict(IN=0, CONST=1, INDEX=2, ADD=8, SUB=9, MUL=10, DIV=11, MOD=12, POW=13, MAX=14, MIN=15, EQ=16, NE=17, LT=18, LE=19,
          GT=20, GE=21, AND=22, OR=23, XOR=24, FMOD=25, FLOORDIV=26, NEG=40, ABS=41, SQRT=42, EXP=43, LOG=44,
          SQUARE=45, RECIP=46, NOT=47, NONZERO=48, ISZERO=49, CAST_F32=56, CAST_I64=57, CAST_I32=58, CAST_BOOL=59,
          CAST_U8=60)


class SpartanError(RuntimeError):
  pass


class sp_program(ctypes.Structure):
  _fields_ = [('n_ops', ctypes.c_int32), ('compute_dtype', ctypes.c_int32),
              ('op', ctypes.c_uint8 * SP_MAX_PROGRAM), ('arg', ctypes.c_uint8 * SP_MAX_PROGRAM),
              ('consts', ctypes.c_double * SP_MAX_CONSTS), ('iconsts', ctypes.c_int64 * SP_MAX_CONSTS),
              ('index_stride', ctypes.c_int64 * 3), ('index_base', ctypes.c_int64)]


class sp_operand(ctypes.Structure):
  _fields_ = [('ptr', ctypes.c_void_p), ('dtype', ctypes.c_int32), ('pad', ctypes.c_int32),
              ('stride', ctypes.c_int64 * 3)]


class sp_gemm_prepared_segment(ctypes.Structure):
  _fields_ = [('A', ctypes.c_void_p), ('B', ctypes.c_void_p), ('Kp', ctypes.c_int64)]


class sp_gemm_prepared_view(ctypes.Structure):
  _fields_ = [('A', ctypes.c_void_p), ('a_copy_stride', ctypes.c_int64), ('B', ctypes.c_void_p),
              ('b_copy_stride', ctypes.c_int64), ('Kp', ctypes.c_int64)]


class sp_gemm_segment(ctypes.Structure):
  _fields_ = [('A', ctypes.c_void_p), ('lda', ctypes.c_int64), ('B', ctypes.c_void_p), ('ldb', ctypes.c_int64),
              ('K', ctypes.c_int64)]


_i64p = ctypes.POINTER(ctypes.c_int64)
_vp, _i64, _int, _dbl, _u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_uint64

_SIGS = {
  'sp_last_error': (ctypes.c_char_p, []),
  'sp_version': (_int, []),
  'sp_init': (_int, [_int]),
  'sp_shutdown': (_int, []),
  'sp_tile_alloc': (_int, [_i64, ctypes.POINTER(_vp), _vp]),
  'sp_tile_free': (_int, [_vp, _vp]),
  'sp_sync': (_int, [_vp]),
  'sp_event_create': (_int, [ctypes.POINTER(_vp)]),
  'sp_event_record': (_int, [_vp, _vp]),
  'sp_event_elapsed': (_int, [_vp, _vp, ctypes.POINTER(ctypes.c_float)]),
  'sp_event_destroy': (_int, [_vp]),
  'sp_device_info': (_int, [ctypes.POINTER(_int), _i64p, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
  'sp_extent_intersection': (_int, [_int, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p]),
  'sp_extent_ravelled_pos': (_i64, [_int, _i64p, _i64p]),
  'sp_extent_unravelled_pos': (_int, [_i64, _int, _i64p, _i64p]),
  'sp_extent_drop_axis': (_int, [_int, _i64p, _i64p, _i64p, _int, _i64p, _i64p, _i64p, ctypes.POINTER(_int)]),
  'sp_extent_to_global': (_i64, [_int, _i64p, _i64p, _i64p, _i64, _int]),
  'sp_extent_change_partition_axis': (_int, [_int, _i64p, _i64p, _i64p, _int, _i64p, _i64p]),
  'sp_good_tile_shape': (_int, [_int, _i64p, _i64, _i64p]),
  'sp_compute_extents': (_i64, [_int, _i64p, _i64p, _i64, _i64p, _i64p, _i64p]),
  'sp_fill': (_int, [_vp, _int, _i64, _int, _dbl, _dbl, _u64, _i64, _vp]),
  'sp_fill2d': (_int, [_vp, _int, _i64, _i64, _i64, _int, _dbl, _dbl, _u64, _i64, _i64, _vp]),
  'sp_map': (_int, [ctypes.POINTER(sp_program), _int, ctypes.POINTER(sp_operand), ctypes.POINTER(sp_operand), _i64p, _vp]),
  'sp_map_reduce_scratch_bytes': (_i64, [_i64p, _int]),
  'sp_map_reduce': (_int, [ctypes.POINTER(sp_program), _int, ctypes.POINTER(sp_operand), ctypes.POINTER(sp_operand),
                           _i64p, _int, _int, _vp, _i64, _vp]),
  'sp_jit_enable': (_int, [_int]),
  'sp_jit_set_nvrtc_path': (_int, [ctypes.c_char_p]),
  'sp_jit_stats': (_int, [_i64p, _i64p, _i64p]),
  'sp_jit_last_log': (ctypes.c_char_p, []),
  'sp_jit_compile_check': (_i64, [ctypes.POINTER(sp_program), _int, _int]),
  'sp_combine': (_int, [_vp, _vp, _int, _i64, _int, _vp]),
  'sp_merge_masked': (_int, [_vp, _i64p, _int, _vp, _i64p, _int, _vp, _i64p, _i64p, _int, _vp]),
  'sp_transpose_2d': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _int, _vp]),
  'sp_copy_rect': (_int, [_vp, _i64p, _vp, _i64p, _i64p, _int, _vp]),
  'sp_gemm_set_chunk_kblocks': (_int, [_int]),
  'sp_gemm_set_variant': (_int, [_int]),
  'sp_gemm_set_round_sync': (_int, [_int]),
  'sp_gemm_set_tuning': (_int, [_int, _int]),
  'sp_upload_2d': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
  'sp_download_2d': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _vp]),
  'sp_gemm_kpad': (_i64, [_i64, _int]),
  'sp_gemm_prepared_bytes': (_i64, [_i64, _i64, _int]),
  'sp_gemm_prepare_a': (_int, [_vp, _i64, _i64, _i64, _int, _vp, _i64, _i64, _i64, _vp]),
  'sp_gemm_prepare_b': (_int, [_vp, _i64, _i64, _i64, _int, _vp, _i64, _i64, _i64, _vp]),
  'sp_gemm_prepared': (_int, [_int, ctypes.POINTER(sp_gemm_prepared_segment), _vp, _i64, _i64, _i64, _int, _int, _vp]),
  'sp_gemm_prepare_a_rows': (_int, [_vp, _i64, _i64, _i64, _int, _vp, _i64, _i64, _i64, _vp]),
  'sp_gemm_prepare_b_rows': (_int, [_vp, _i64, _i64, _i64, _int, _vp, _i64, _i64, _i64, _vp]),
  'sp_gemm_prepared_views': (_int, [_int, ctypes.POINTER(sp_gemm_prepared_view), _vp, _i64, _i64, _i64, _int, _int, _vp]),
  'sp_gemm_prepared_views_gated': (_int, [_int, ctypes.POINTER(sp_gemm_prepared_view), ctypes.POINTER(_vp),
                                          ctypes.POINTER(ctypes.c_uint32), _vp, _vp, _i64, _i64, _i64, _int, _int, _vp]),
  'sp_peer_alloc': (_int, [_i64, ctypes.POINTER(_vp)]),
  'sp_peer_free': (_int, [_vp]),
  'sp_peer_handle_bytes': (_int, []),
  'sp_peer_export': (_int, [_vp, _vp]),
  'sp_peer_import': (_int, [_vp, ctypes.POINTER(_vp)]),
  'sp_peer_close': (_int, [_vp]),
  'sp_peer_push': (_int, [_int, ctypes.POINTER(_vp), _vp, _i64, ctypes.POINTER(_vp), _vp, _vp]),
  'sp_peer_push_2d': (_int, [_int, ctypes.POINTER(_vp), _i64, _vp, _i64, _i64, _i64, ctypes.POINTER(_vp), _vp, _vp]),
  'sp_write_u32': (_int, [_vp, ctypes.c_uint32, _vp]),
  'sp_wait_u32': (_int, [_vp, ctypes.c_uint32, _i64, _vp, _vp]),
  'sp_gemm_prepared_kmeans': (_int, [ctypes.POINTER(sp_gemm_prepared_segment), _i64, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _int, _vp]),
  'sp_gemm_argmin_parts': (_i64, [_i64]),
  'sp_gemm_prepared_argmin': (_int, [_int, ctypes.POINTER(sp_gemm_prepared_segment), _i64, _i64, _vp, _vp, _vp, _int, _vp]),
  'sp_gemm_f32_workspace_bytes': (_i64, [_i64, _i64, _int, _i64p, _int]),
  'sp_gemm_f32_segments': (_int, [_int, ctypes.POINTER(sp_gemm_segment), _vp, _i64, _i64, _i64, _int, _int, _vp, _i64, _vp]),
  'sp_gemm_f32': (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _i64, _vp]),
  'sp_gemm_f32_ex': (_int, [_vp, _i64, _int, _vp, _i64, _int, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _i64, _vp]),
  'sp_kmeans_workspace_bytes': (_i64, [_i64, _i64, _i64]),
  'sp_kmeans_assign': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
  'sp_kmeans_prepared_bytes': (_i64, [_i64, _i64]),
  'sp_kmeans_prepare_points': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp]),
  'sp_kmeans_set_fused': (_int, [_int]),
  'sp_kmeans_assign_workspace_bytes': (_i64, [_i64, _i64, _i64]),
  'sp_kmeans_assign_prepared': (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
  'sp_spmv_csr': (_int, [_vp, _int, _vp, _vp, _i64, _vp, _vp, _int, _int, _vp]),
  'sp_gemm_simt': (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp]),
}

EXPORTS = sorted(_SIGS)

for _name, (_res, _args) in _SIGS.items():
  _fn = getattr(lib, _name)     # AttributeError here = the .so does not export what the header declares
  _fn.restype = _res
  _fn.argtypes = _args


def last_error():
  return lib.sp_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
  """Raise on a negative sp_status."""
  if rc is not None and rc < 0:
    raise SpartanError('%s failed (sp_status %d): %s' % (what or 'libspartan_b200 call', rc, last_error()))
  return rc


def i64arr(values):
  values = [int(v) for v in values]
  return (ctypes.c_int64 * max(1, len(values)))(*values)
