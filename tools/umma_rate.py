"""tcgen05 dispatch-rate microbenchmark (libg4d_selftest.so): cycles per tcgen05.mma as a function of N, operand source and
concurrent TMEM<->register traffic.  Output feeds DESIGN.md's pipe model of the geometry kernel."""
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
lib.g4d_selftest_umma_rate.argtypes = [C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p]
lib.g4d_selftest_umma_rate.restype = C.c_int
ND = 2048
rows = []
for blocks in (1, 148):
    for (kind, a_smem) in ((0, 0), (0, 1), (1, 0)):
        for N in (48, 64, 128, 256):
            for (readers, op) in ((0, 0), (4, 1), (8, 1), (8, 2)):
                for apart in ((0, 128) if N <= 128 else (0,)):
                    if blocks == 148 and (apart or readers == 4):
                        continue
                    out = torch.zeros(blocks * 4, dtype=torch.int64, device="cuda")
                    cfg = (C.c_int * 8)(N, a_smem, kind, readers, op, ND, apart, 0)
                    rc = lib.g4d_selftest_umma_rate(cfg, out.data_ptr(), blocks, None)
                    torch.cuda.synchronize()
                    o = out.view(blocks, 4).double().mean(0).tolist()
                    rows.append({"blocks": blocks, "kind": "tf32" if kind else "f16", "A": "smem" if a_smem else "tmem", "N": N,
                                 "readers": readers, "op": ["-", "ld", "st"][op], "acc_apart": apart, "rc": rc,
                                 "cyc_per_disp": round(o[1] / ND, 1), "issue_cyc_per_disp": round(o[0] / ND, 1),
                                 "reader_B_per_cyc": round(o[2] * 2048 / max(o[1], 1), 1)})
                    print(json.dumps(rows[-1]), flush=True)
