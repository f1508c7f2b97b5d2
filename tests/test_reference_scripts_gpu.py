"""The reference's UNMODIFIED training scripts (arxiv_pyg/gnn.py, gnn_kd_and_aux.py) executed end to end on a GPU on
top of the shim packages — BASELINE.json north_star: "gnn.py and gnn_kd_and_aux.py run unmodified".

The scripts are not part of this repository: `__graft_entry__.build()` stages them, byte for byte, from /root/reference
into baseline/_ref/ (git-ignored) when the reference tree is present; the test is skipped when neither location exists.
Each run happens in a scratch directory laid out like the reference checkout (arxiv_pyg/ beside arxiv_dgl/ with the
teacher artefacts `logits/<expt>/<seed>.pt`, `features/<expt>/<seed>.pt` the scripts torch.load — arxiv_pyg/gnn.py:274-275),
with PYTHONPATH pointing at the shims.  Evidence that the b200gnn kernels did the work: the launch counter of
libb200gnn.so, written at interpreter exit (B200GNN_LAUNCH_REPORT)."""
import filecmp
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = Path(__file__).resolve().parents[1]
CANDIDATES = [Path("/root/reference/arxiv_pyg"), ROOT / "baseline" / "_ref" / "arxiv_pyg"]
FILES = ["gnn.py", "gnn_kd_and_aux.py", "criterion.py", "logger.py"]


def _source_dir():
    for c in CANDIDATES:
        if (c / "gnn.py").exists():
            return c
    return None


@pytest.fixture(scope="module")
def checkout(tmp_path_factory):
    src = _source_dir()
    if src is None:
        pytest.skip("reference scripts not staged (run __graft_entry__.build() where /root/reference exists)")
    base = tmp_path_factory.mktemp("refrun")
    wd = base / "arxiv_pyg"
    wd.mkdir()
    for f in FILES:
        shutil.copyfile(src / f, wd / f)
        assert filecmp.cmp(src / f, wd / f, shallow=False)
    # teacher artefacts in the reference's format (arxiv_dgl/gat.py:247-251 writes them, gnn.py:274-275 reads them)
    sys.path.insert(0, str(ROOT))
    import efficient_gnns_b200  # noqa: F401
    from efficient_gnns_b200 import synthetic
    ds = synthetic.make_node_dataset(synthetic.ARXIV, seed=0)
    for kind, t in (("logits", ds.teacher_logits), ("features", ds.teacher_feat)):
        d = base / "arxiv_dgl" / kind / "gat-3L250x3h"
        d.mkdir(parents=True)
        torch.save(t, d / "0.pt")
    return wd


def _run(wd, script, *args):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(ROOT / "efficient-gnns_b200" / "shim"), str(ROOT), env.get("PYTHONPATH", "")])
    report = wd / f"launches_{script}_{'_'.join(a.strip('-') for a in args[:2])}.txt"
    env["B200GNN_LAUNCH_REPORT"] = str(report)
    p = subprocess.run([sys.executable, script, "--runs", "1", "--epochs", "2", "--expt_name", "b200gnn", *args], cwd=wd, env=env,
                       capture_output=True, text=True, timeout=800)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    m = re.findall(r"Epoch: (\d+), Loss_total: ([-0-9.naninf]+), Loss_cls: ([-0-9.naninf]+)", p.stdout)
    assert [int(e) for e, _, _ in m] == [1, 2], p.stdout[-2000:]
    losses = [float(l) for _, l, _ in m]
    assert all(l == l and abs(l) < 1e6 for l in losses), losses
    assert int(report.read_text()) > 20, "libb200gnn.so kernels did not launch"
    res = list((wd / "logs").rglob("results.pt"))
    assert res, "the script did not reach its torch.save(results.pt)"
    return p.stdout, losses


@pytest.mark.parametrize("gnn", ["gcn", "sage"])
def test_gnn_py_supervised_runs_unmodified(checkout, gnn):
    out, losses = _run(checkout, "gnn.py", "--gnn", gnn, "--training", "supervised")
    assert 2.0 < losses[0] < 6.0          # ~ln(40) for 40 random classes at initialisation
    assert "Highest Train" in out or "Final Train" in out


def test_gnn_py_kd_reads_teacher_artefacts(checkout):
    _run(checkout, "gnn.py", "--gnn", "gcn", "--training", "kd")


def test_gnn_py_lpw_uses_subgraph_and_segment_softmax(checkout):
    _run(checkout, "gnn.py", "--gnn", "gcn", "--training", "lpw", "--kernel", "cosine")


def test_gnn_kd_and_aux_py_nce(checkout):
    _run(checkout, "gnn_kd_and_aux.py", "--gnn", "gcn", "--training", "nce", "--max_samples", "4096")
