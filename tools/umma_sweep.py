"""Sweep the tcgen05 descriptor / layout hypotheses of g4d_debug_umma on a real B200 (each case in its own
subprocess with a timeout so that a wrong encoding cannot hang or poison the rest)."""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(cfg):
    import torch
    g4d = importlib.import_module("4dgaussians_b200")
    lib = g4d._lib.load_selftest()
    ws = g4d._lib.Workspace.get(0)
    N, K = cfg[0], cfg[1]
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(128, K, generator=g)).cuda()
    B = (torch.randn(N, K, generator=g)).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    arr = (C.c_int * 8)(*cfg)
    rc = lib.g4d_selftest_umma(arr, A.data_ptr(), B.data_ptr(), D.data_ptr(), 0)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t())
    err = (D.double() - ref).abs().max().item()
    # also: error against the transposed / permuted candidates to help diagnose layouts
    print(json.dumps({"cfg": cfg, "rc": rc, "max_abs_err": err, "ref_max": ref.abs().max().item(),
                      "nan": bool(torch.isnan(D).any().item())}))


def main():
    if len(sys.argv) > 1:
        one(json.loads(sys.argv[1]))
        return
    cases = []
    for (N, K) in ((128, 128), (128, 32), (48, 64), (16, 64)):
        for layout_mode in (0, 1):
            for swap in (0, 1):
                for tma in (0, 1):
                    if tma and layout_mode == 1:
                        continue
                    cases.append([N, K, layout_mode, swap, 1, tma, 0, 1])
    cases.append([128, 128, 0, 0, 1, 0, 1, 1])   # single-pass TF32 (expected err ~1e-2)
    cases.append([128, 128, 0, 0, 1, 0, 0, 0])   # version bit off
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(c)], capture_output=True, text=True, timeout=60)
            out = (r.stdout.strip().splitlines() or ["<no output> rc=%d %s" % (r.returncode, r.stderr.strip()[-300:])])[-1]
        except subprocess.TimeoutExpired:
            out = json.dumps({"cfg": c, "TIMEOUT": True})
        print(out, flush=True)


if __name__ == "__main__":
    main()
