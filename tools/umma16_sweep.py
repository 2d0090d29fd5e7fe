"""bf16 (kind::f16) tcgen05 self-test sweep: operand roles used by the backward kernels."""
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
ws = g4d._lib.Workspace.get(0)
cases = []
for (N, K) in ((128, 128), (48, 64), (128, 48), (16, 128)):
    for a_mode, pack in ((0, 0), (0, 1), (1, 0), (2, 0)):
        for b_mode in (0, 1):
            cases.append([N, K, a_mode, b_mode, pack, 0, 16, 0])
cases.append([128, 128, 1, 0, 0, 1, 16, 0])     # single-pass bf16
only = json.loads(sys.argv[1]) if len(sys.argv) > 1 else None
for c in (only or cases):
    N, K = c[0], c[1]
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, K, generator=g).cuda(); B = torch.randn(N, K, generator=g).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    try:
        rc = lib.g4d_selftest_umma((C.c_int * 8)(*c), A.data_ptr(), B.data_ptr(), D.data_ptr(), 0)
        torch.cuda.synchronize()
        ref = A.double() @ B.double().t()
        err = (D.double() - ref).abs().max().item() / ref.abs().max().item()
        print(json.dumps({"cfg": c, "rc": rc, "rel_err": err}), flush=True)
    except Exception as e:
        print(json.dumps({"cfg": c, "EXC": str(e)[:200]}), flush=True)
        break
