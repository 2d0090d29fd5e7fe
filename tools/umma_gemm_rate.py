"""tcgen05 microbenchmark, realistic variant: cycles per [128 x N x 128] 3-product GEMM of the deformation MLP (operand walk of the
forward kernels), f16 vs tf32 operands, A in TMEM vs shared memory, with / without concurrent TMEM<->register traffic."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

g4d = importlib.import_module("4dgaussians_b200")
lib = g4d._lib.load_selftest()
lib.g4d_selftest_umma_gemm_rate.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p]
lib.g4d_selftest_umma_gemm_rate.restype = C.c_int
NG = 64
for blocks in (1, 148):
    for (kind, a_smem) in ((0, 0), (0, 1), (1, 0)):
        for (N, split) in ((128, 1), (128, 2), (64, 1)):
            for distinct in (1, 3):      # bit 1: SWIZZLE_128B images
                for (readers, op) in ((0, 0), (8, 1)):
                    if blocks == 148 and op:
                        continue
                    out = torch.zeros(blocks * 4, dtype=torch.int64, device="cuda")
                    cfg = (C.c_int * 8)(N, a_smem, kind, readers, op, NG, distinct, split)
                    rc = lib.g4d_selftest_umma_gemm_rate(cfg, out.data_ptr(), blocks, None)
                    torch.cuda.synchronize()
                    o = out.view(blocks, 4).double().mean(0).tolist()
                    nd = NG * split * 3 * (16 if kind else 8)
                    print(json.dumps({"blocks": blocks, "kind": "tf32" if kind else "f16", "A": "smem" if a_smem else "tmem", "N": N,
                                      "split": split, "swz128": distinct >> 1, "readers": readers, "op": ["-", "ld", "st"][op], "rc": rc,
                                      "cyc_per_gemm": round(o[1] / NG, 1), "cyc_per_disp": round(o[1] / nd, 1),
                                      "reader_B_per_cyc": round(o[2] * 2048 / max(o[1], 1), 1)}), flush=True)
