"""Does the tcgen05 fp32 accumulator round to nearest or truncate?  Signed error of the 3xTF32 GEMMs against fp64 for
(a) random-sign operands and (b) all-positive operands (a truncating accumulator shows up as a negative mean error
that grows with the number of accumulation steps), next to torch.mm fp32 (cuBLAS) on the same inputs."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa: E402,F401
from efficient_gnns_b200 import ops  # noqa: E402


def stats(c, ref):
    e = (c.double() - ref) / ref.abs().max()
    er = ((c.double() - ref) / ref.abs().clamp_min(1e-300))
    return dict(max_rel=float(e.abs().max()), fro_rel=float((c.double() - ref).norm() / ref.norm()),
                mean_signed_rel=float(er.mean()))


def main():
    torch.manual_seed(0)
    dev = "cuda"
    for positive in (False, True):
        for K in (32, 128, 256, 1024):
            M, N = 4096, 256
            a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev)
            if positive:
                a, b = a.abs(), b.abs()
            ref = a.double() @ b.double().t()
            hi, lo = ops.split_tf32(b)
            c = ops.gemm_tf32x3(a, hi, lo)
            c32 = a @ b.t()
            print(json.dumps(dict(op="gemm", positive=positive, K=K, tf32x3=stats(c, ref), cublas_fp32=stats(c32, ref))), flush=True)
        for rows in (148 * 16, 148 * 16 * 8, 148 * 16 * 72):
            x = torch.randn(rows, 256, device=dev); g = torch.randn(rows, 256, device=dev)
            if positive:
                x, g = x.abs(), g.abs()
            ref = x.double().t() @ g.double()
            w = ops.gemm_wgrad_tf32x3(x, g)
            w32 = x.t() @ g
            print(json.dumps(dict(op="wgrad", positive=positive, rows=rows, chain_mmas=rows // 148 // 8 * 3, tf32x3=stats(w, ref),
                                  cublas_fp32=stats(w32, ref))), flush=True)


if __name__ == "__main__":
    main()
