"""A/B of the tensor-core GEMM variants on one box: one CTA per 128x256 tile vs CTA pairs (cta_group::2).
Times sp_gemm_prepared alone (operands prepared once), interleaving the variants; prints ms and algorithmic TFLOP/s."""
import os, sys, json, subprocess
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import spartan_b200 as sp
from spartan_b200 import device_ops
from spartan_b200._lib import lib, check

ctx = sp.initialize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
precs = sys.argv[2].split(',') if len(sys.argv) > 2 else ['bf16x3']
chunks = [int(c) for c in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]
A = torch.rand(n, n, device='cuda'); B = torch.rand(n, n, device='cuda'); C = torch.empty(n, n, device='cuda')

def clocks():
  try:
    out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-i', '0'],
                         capture_output=True, text=True, timeout=5).stdout.strip()
    return out
  except Exception:
    return '?'

rows = []
for prec in precs:
  Kp = device_ops.gemm_kpad(n, prec)
  pa = torch.zeros(device_ops.gemm_prepared_bytes(n, Kp, prec), dtype=torch.uint8, device='cuda')
  pb = torch.zeros(device_ops.gemm_prepared_bytes(n, Kp, prec), dtype=torch.uint8, device='cuda')
  device_ops.gemm_prepare_a(A, pa, Kp, 0, prec); device_ops.gemm_prepare_b(B, pb, Kp, 0, prec)
  for chunk in chunks:
    check(lib.sp_gemm_set_chunk_kblocks(chunk), 'chunk')
    for rep in range(3):
      for variant in (1, 2):
        check(lib.sp_gemm_set_variant(variant), 'variant')
        device_ops.gemm_prepared([(pa, pb, Kp)], C, False, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps): device_ops.gemm_prepared([(pa, pb, Kp)], C, False, prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row = {'n': n, 'prec': prec, 'chunk_kb': chunk, 'variant': variant, 'ms': round(ms, 3),
               'tflops': round(2 * n ** 3 / ms / 1e9, 1), 'clk_pwr_after': clocks()}
        print(json.dumps(row), flush=True); rows.append(row)
check(lib.sp_gemm_set_variant(0), 'variant'); check(lib.sp_gemm_set_chunk_kblocks(0), 'chunk')
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'gemm_variant_probe.json'), 'w'), indent=1)
