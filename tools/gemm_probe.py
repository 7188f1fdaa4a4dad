"""GPU bring-up probe for the tcgen05 GEMM (run under gpurun; not a test, not a bench).

Checks small/ragged shapes against an fp64 product, measures the accumulation error of a
long-K dot with operands that are exactly representable in TF32 (isolates the tensor-core
accumulator rounding), and times large squares.  Prints one JSON line per experiment.
"""
import ctypes
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "..", "spartan_b200", "libspartan_b200.so"))
lib.sp_last_error.restype = ctypes.c_char_p
lib.sp_gemm_f32_workspace_bytes.restype = ctypes.c_int64
lib.sp_gemm_f32_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
lib.sp_gemm_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                            ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
lib.sp_gemm_set_chunk_kblocks.argtypes = [ctypes.c_int]
lib.sp_gemm_simt.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                             ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                             ctypes.c_int, ctypes.c_void_p]

_ws = {}


def gemm(A, B, C, precision, accumulate=0):
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    ks = (ctypes.c_int64 * 1)(K)
    need = lib.sp_gemm_f32_workspace_bytes(M, N, 1, ks, precision)
    ws = _ws.get("ws")
    if ws is None or ws.numel() < need:
        _ws["ws"] = None
        ws = torch.empty(need, dtype=torch.uint8, device="cuda")
        _ws["ws"] = ws
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.sp_gemm_f32(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), C.stride(0), M, N, K,
                         accumulate, precision, ws.data_ptr(), ws.numel(), stream)
    if rc != 0:
        raise RuntimeError("sp_gemm_f32 rc=%d: %s" % (rc, lib.sp_last_error().decode()))


def errs(C, ref):
    d = (C.double() - ref).abs()
    return {"max_abs": d.max().item(), "max_rel_to_max": (d.max() / ref.abs().max()).item(),
            "max_elem_rel": (d / ref.abs().clamp_min(1e-30)).max().item(),
            "mean_signed_rel": ((C.double() - ref) / ref.abs().clamp_min(1e-30)).mean().item()}


MODES = [(0, "tf32x1"), (1, "tf32x3"), (3, "bf16x3")]


def check_shapes():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 32), (128, 256, 64), (256, 512, 128), (100, 77, 50), (132, 77, 100),
                      (300, 700, 1000), (1024, 1024, 1024), (4096, 4096, 4096)]:
        A = torch.rand(M, K, device="cuda")
        B = torch.rand(K, N, device="cuda")
        ref = A.double() @ B.double()
        for prec, name in MODES:
            C = torch.full((M, N), float("nan"), device="cuda")
            gemm(A, B, C, prec)
            torch.cuda.synchronize()
            e = errs(C, ref)
            print(json.dumps({"exp": "shape", "M": M, "N": N, "K": K, "prec": name, **e}), flush=True)
        # accumulate flag
        C = torch.ones(M, N, device="cuda")
        gemm(A, B, C, 1, accumulate=1)
        torch.cuda.synchronize()
        e = errs(C, ref + 1.0)
        print(json.dumps({"exp": "accumulate", "M": M, "N": N, "K": K, **e}), flush=True)
        # simt cross-check in fp64
        Ad, Bd = A.double(), B.double()
        Cd = torch.empty(M, N, device="cuda", dtype=torch.float64)
        rc = lib.sp_gemm_simt(Ad.data_ptr(), K, Bd.data_ptr(), N, Cd.data_ptr(), N, M, N, K, 1, 0,
                              torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.sp_last_error()
        torch.cuda.synchronize()
        print(json.dumps({"exp": "simt_f64", "M": M, "N": N, "K": K, **errs(Cd, ref)}), flush=True)


def accum_precision():
    # operands exactly representable in tf32 AND bf16-split -> products exact; any error is accumulation.
    torch.manual_seed(1)
    M, N = 256, 512
    for K in [4096, 32768]:
        A = (torch.randint(0, 256, (M, K), device="cuda").float() / 256.0)
        B = (torch.randint(0, 256, (K, N), device="cuda").float() / 256.0)
        ref = A.double() @ B.double()
        out = {"exp": "accum", "K": K}
        for prec, name in MODES:
            for chunk in [0, 1, 2, 4, 8, 16, 1 << 20]:
                lib.sp_gemm_set_chunk_kblocks(chunk)
                C = torch.empty(M, N, device="cuda")
                gemm(A, B, C, prec)
                torch.cuda.synchronize()
                e = errs(C, ref)
                out["%s_chunk%d" % (name, chunk)] = [e["max_rel_to_max"], e["mean_signed_rel"]]
        lib.sp_gemm_set_chunk_kblocks(0)
        out["torch_fp32"] = errs(A @ B, ref)["max_rel_to_max"]
        print(json.dumps(out), flush=True)
    for K in [4096, 32768]:
        for gen, gname in [(torch.randn, "randn"), (torch.rand, "rand")]:
            A = gen(M, K, device="cuda")
            B = gen(K, N, device="cuda")
            ref = A.double() @ B.double()
            out = {"exp": gname, "K": K}
            for prec, name in MODES:
                C = torch.empty(M, N, device="cuda")
                gemm(A, B, C, prec)
                torch.cuda.synchronize()
                e = errs(C, ref)
                out[name] = [e["max_rel_to_max"], e["max_elem_rel"], e["mean_signed_rel"]]
            e = errs(A @ B, ref)
            out["torch_fp32"] = [e["max_rel_to_max"], e["max_elem_rel"], e["mean_signed_rel"]]
            print(json.dumps(out), flush=True)


def timing():
    for n in [8192, 16384, 32768]:
        A = torch.rand(n, n, device="cuda")
        B = torch.rand(n, n, device="cuda")
        C = torch.empty(n, n, device="cuda")
        for prec, name in MODES:
            for chunk in ([0, 8, 1 << 20] if n == 16384 else [0]):
                lib.sp_gemm_set_chunk_kblocks(chunk)
                reps = 1 if n == 32768 else 3
                gemm(A, B, C, prec)
                torch.cuda.synchronize()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                for _ in range(reps):
                    gemm(A, B, C, prec)
                ev1.record()
                torch.cuda.synchronize()
                ms = ev0.elapsed_time(ev1) / reps
                ref = A[:128].double() @ B[:, :256].double()
                err = errs(C[:128, :256], ref)["max_rel_to_max"]
                print(json.dumps({"exp": "time", "n": n, "prec": name, "chunk": chunk, "ms": ms,
                                  "tflops": 2.0 * n ** 3 / ms / 1e9, "max_rel_to_max": err}), flush=True)
        lib.sp_gemm_set_chunk_kblocks(0)
        del A, B, C
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["shapes", "accum", "time"]
    print(json.dumps({"device": torch.cuda.get_device_name(0)}), flush=True)
    if "shapes" in which:
        check_shapes()
    if "accum" in which:
        accum_precision()
    if "time" in which:
        timing()
