"""Numerical sensitivity probe: the 1-GPU engine on a graph and on a random relabelling of the same graph
(same mathematics, different summation orders), both against the fp64 CPU oracle."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import efficient_gnns_b200  # noqa
from efficient_gnns_b200 import sparse, synthetic
from efficient_gnns_b200.engine import GCNStudentTrainer
from oracle import graph as og, nn as onn, criterion as oc

n, e, dims = 20_011, 150_000, [128, 256, 256, 40]
ei = synthetic.skewed_edges(n, e, 0)
g = torch.Generator().manual_seed(1)
x = torch.randn(n, dims[0], generator=g); y = torch.randint(0, dims[-1], (n,), generator=g)
t = torch.randn(n, dims[-1], generator=g) * 2
idx = torch.randperm(n, generator=g)[: n // 2].sort().values

def build(ei_, tc=True):
    d = ei_.cuda()
    perm = (d[1] * n + d[0]).argsort()
    adj = sparse.SparseTensor(row=d[1][perm], col=d[0][perm], sparse_sizes=(n, n), is_sorted=True).to_symmetric()
    return GCNStudentTrainer(adj, dims, dropout=0.0, seed=3, tensor_core_gemm=tc)

def grads(tr):
    out = {}
    for l in range(tr.L):
        out[f"W{l}"] = tr.gW[l].double().cpu()
        if l < tr.L - 1:
            out[f"gamma{l}"] = tr.ggamma[l].double().cpu(); out[f"beta{l}"] = tr.gbeta[l].double().cpu()
    out[f"b{tr.L-1}"] = tr.gb[-1].double().cpu()
    return out

# fp64 oracle
row, col, _ = og.to_sparse_adj_t(ei.numpy(), n); r, c = og.to_symmetric(row, col, n); r, c, v = og.gcn_norm(r, c, n)
ptr, c, v = torch.from_numpy(og.ind2ptr(r, n)), torch.from_numpy(c), torch.from_numpy(v).double()
a = build(ei)
sd = {k: w.double().cpu() for k, w in a.state_dict().items()}
W = [sd[f"convs.{i}.weight"].clone().requires_grad_(True) for i in range(3)]
Bb = [sd[f"convs.{i}.bias"].clone().requires_grad_(True) for i in range(3)]
ga = [sd[f"bns.{i}.weight"].clone().requires_grad_(True) for i in range(2)]
be = [sd[f"bns.{i}.bias"].clone().requires_grad_(True) for i in range(2)]
logits, _ = onn.gcn_forward(x.double(), ptr, c, v, W, Bb, ga, be, None)
loss, _, _ = oc.kd_criterion(logits[idx], y[idx], t.double()[idx], 0.9, 4.0)
loss.backward()
ref = {f"W{i}": W[i].grad for i in range(3)}; ref.update({f"gamma{i}": ga[i].grad for i in range(2)})
ref.update({f"beta{i}": be[i].grad for i in range(2)}); ref["b2"] = Bb[2].grad

a.train_step(x.cuda(), y.cuda(), idx.cuda(), t.cuda()); ga_ = grads(a)
# relabelled copy
p = torch.randperm(n, generator=g); inv = torch.empty(n, dtype=torch.long); inv[p] = torch.arange(n)
b = build(inv[ei]); b.train_step(x[p].cuda(), y[p].cuda(), inv[idx].sort().values.cuda(), t[p].cuda()); gb_ = grads(b)
c_ = build(ei, tc=False); c_.train_step(x.cuda(), y.cuda(), idx.cuda(), t.cuda()); gc_ = grads(c_)
print(f"{'tensor':8s} {'|ref|max':>10s} {'engine-ref':>11s} {'perm-ref':>11s} {'engine-perm':>11s} {'cublas-ref':>11s}")
for k in ref:
    m = ref[k].abs().max().item()
    print(f"{k:8s} {m:10.3e} {(ga_[k]-ref[k]).abs().max().item()/m:11.2e} {(gb_[k]-ref[k]).abs().max().item()/m:11.2e} "
          f"{(ga_[k]-gb_[k]).abs().max().item()/m:11.2e} {(gc_[k]-ref[k]).abs().max().item()/m:11.2e}")

# ---- isolate the tensor-core kernels on the engine's own buffers (engine `a`, state after one step)
def rel(x_, r_):
    return ((x_.double().cpu() - r_).abs().max() / r_.abs().max()).item()
a2 = build(ei)
a2.forward(x.cuda(), training=True)
a2.dY[-1].zero_()
from efficient_gnns_b200 import ops
ops.kd_loss_fwd_bwd(a2.Y[-1], y.cuda(), idx.cuda(), t.cuda(), 0.9, 4.0, d_logits=a2.dY[-1], loss_out=a2.loss_out, partial=a2.kd_part)
a2.backward(x.cuda())
torch.cuda.synchronize()
d = lambda z: z.double().cpu()
print("wgrad W1   : tc vs fp64(A0^T dH1)", rel(a2.gW[1], d(a2.A[0]).t() @ d(a2.dH[1])))
print("wgrad W0   : tc vs fp64(AX^T dY0)", rel(a2.gW[0], d(a2.AX).t() @ d(a2.dY[0])))
print("wgrad W2   : tc vs fp64(A1^T dH2)", rel(a2.gW[2], d(a2.A[1]).t() @ d(a2.dH[2])))
print("dgrad dA1  : tc vs fp64(dH2 W2^T)", rel(a2.dA[1], d(a2.dH[2]) @ d(a2.W[2]).t()))
print("dgrad dA0  : tc vs fp64(dH1 W1^T)", rel(a2.dA[0], d(a2.dH[1]) @ d(a2.W[1]).t()))
print("fwd H1     : tc vs fp64(A0 W1)   ", rel(a2.H[1], d(a2.A[0]) @ d(a2.W[1])))
print("fwd Y0     : tc vs fp64(AX W0+b) ", rel(a2.Y[0], d(a2.AX) @ d(a2.W[0]) + d(a2.b[0])))
print("colsum dA1 : tc", (a2.dA[1].double().sum(0).cpu() - (d(a2.dH[2]) @ d(a2.W[2]).t()).sum(0)).abs().max().item(),
      "of", (d(a2.dH[2]) @ d(a2.W[2]).t()).sum(0).abs().max().item())
