import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa
from efficient_gnns_b200 import ops

torch.set_printoptions(precision=4, linewidth=200)
for (Nn, Kin, Nout) in [(16, 128, 32), (16, 128, 256), (32, 256, 64)]:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Nn, Kin, generator=g); d = torch.randn(Nn, Nout, generator=g)
    ref = x.double().t() @ d.double()
    out = ops.gemm_wgrad_tf32x3(x.cuda(), d.cuda()).cpu().double()
    print("shape", Nn, Kin, Nout, "nonzero frac", (out != 0).float().mean().item(), "nan", torch.isnan(out).any().item())
    print("ref[:3,:6]", ref[:3, :6]); print("out[:3,:6]", out[:3, :6])
    r = out / ref
    print("ratio median", r.median().item(), "max abs err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item())
    # one-hot probes: where does X[k0,m0]*G[k0,n0] land?
    for (k0, m0, n0) in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (3, 5, 7), (9, 40, 20), (12, 100, 31)]:
        if n0 >= Nout: continue
        x = torch.zeros(Nn, Kin); d = torch.zeros(Nn, Nout); x[k0, m0] = 1.0; d[k0, n0] = 1.0
        o = ops.gemm_wgrad_tf32x3(x.cuda(), d.cuda()).cpu()
        nz = o.nonzero().tolist()
        print("probe", (k0, m0, n0), "->", [(i, j, round(o[i, j].item(), 4)) for i, j in nz[:8]], "count", len(nz))
