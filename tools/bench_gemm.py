"""Micro-benchmark: tcgen05 3xTF32 GEMM vs cuBLAS fp32 (torch.mm) on the layer shapes of the ARXIV-shape GCN."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import ops  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    M = 169_343
    for (K, N) in [(128, 256), (256, 256), (256, 40), (40, 256)]:
        a = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") / K ** 0.5
        hi, lo = ops.split_tf32(w)
        out = torch.empty(M, N, device="cuda")
        t_tc = timeit(lambda: ops.gemm_tf32x3(a, hi, lo, out=out))
        wt = w.t().contiguous()
        t_cb = timeit(lambda: torch.mm(a, wt, out=out))
        ref = (a[:4096].double() @ w.double().t())
        err = ((out[:4096].double() - ref).abs().max() / ref.abs().max()).item()
        flops = 2.0 * M * N * K
        print(json.dumps(dict(M=M, N=N, K=K, ms_tcgen05=t_tc, ms_cublas_fp32=t_cb, tflops_effective=flops / t_tc / 1e9,
                              hbm_GBps=(M * K + M * N) * 4 / t_tc / 1e6, rel_err=err)), flush=True)


def wgrad():
    M = 169_343
    for (K, N) in [(128, 256), (256, 256)]:
        x = torch.randn(M, K, device="cuda"); d = torch.randn(M, N, device="cuda")
        out = torch.empty(K, N, device="cuda")
        ws = torch.empty(148 * K * N, device="cuda")
        t_tc = timeit(lambda: ops.gemm_wgrad_tf32x3(x, d, out=out, workspace=ws))
        t_cb = timeit(lambda: torch.mm(x.t(), d, out=out))
        print(json.dumps(dict(kind="wgrad", Nn=M, Kin=K, Nout=N, ms_tcgen05=t_tc, ms_cublas_fp32=t_cb,
                              hbm_GBps=(M * K + M * N) * 4 / t_tc / 1e6)), flush=True)


if __name__ == "__main__":
    wgrad()
    main()
