"""A/B of the pair GEMM with and without the per-round soft barrier at the benchmark size: sustained runs
(5 launches back to back per sample, interleaved), ms / TFLOP/s / SM clock, and bit-equality of the results."""
import os, sys, json, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200 import device_ops
from spartan_b200._lib import lib, check
ctx = sp.initialize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
prec = 'bf16x3'
A = torch.rand(n, n, device='cuda'); B = torch.rand(n, n, device='cuda'); C = torch.empty(n, n, device='cuda')
pa = device_ops.PreparedOperand(n, n, prec, 'pa'); pb = device_ops.PreparedOperand(n, n, prec, 'pb')
pa.prepare_a(A, 0); pb.prepare_b(B, 0)
del A, B
def clocks():
  out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-i', '0'],
                       capture_output=True, text=True, timeout=5).stdout.strip()
  return out
ref = None
CONFIGS = [(0, 0, 16), (1, 0, 16), (1, 0, 8), (1, 0, 4), (1, 128, 16), (1, 128, 8), (1, 32, 8)]
for rep in range(2):
  for on, sync_kb, group_m in CONFIGS:
    check(lib.sp_gemm_set_round_sync(on), 'sync'); check(lib.sp_gemm_set_tuning(sync_kb, group_m), 'tuning')
    device_ops.gemm_prepared_rows(pa, 0, n, pb, 0, n, C)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): device_ops.gemm_prepared_rows(pa, 0, n, pb, 0, n, C)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    chk = float(C[::997, ::991].double().sum())
    if ref is None: ref = chk
    print(json.dumps({'round_sync': on, 'sync_kb': sync_kb, 'group_m': group_m, 'ms': round(ms, 2),
                      'tflops': round(2 * n ** 3 / ms / 1e9, 1), 'clk_pwr': clocks(), 'same_bits': chk == ref}), flush=True)
check(lib.sp_gemm_set_round_sync(1), "sync"); check(lib.sp_gemm_set_tuning(0, 8), "tuning")
