"""Multi-rank correctness on real GPUs (skipped when fewer than 2 are visible): the two places the path has a collective.
 * config 5: the block-cyclic Cholesky over 2 ranks (NCCL panel broadcasts, look-ahead) must be BIT-IDENTICAL to the
   single-GPU factor of lb_fit (same per-tile update order): max |L - L_single| == 0.0;
 * candidate sharding: sharded argmax (one all_gather) == unsharded device argmax, ties to the lowest global index.
The rank logic is also covered on CPU with gloo (tests/test_dist_gloo.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("size", [1000, 8192])
def test_two_rank_cholesky_is_bit_identical_to_single_gpu(size):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun(2, "tools/dist_chol_run.py", "--size", str(size), "--steps", "1", "--check", "gather")
    assert res["n_gpus"] == 2 and res["info"] == 0
    assert res["max_abs_diff_vs_single_gpu"] == 0.0, res
    assert res["logdet_rel_diff"] <= 1e-13, res


def test_two_rank_sharded_argmax_matches_unsharded():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun(2, "tools/dist_argmax_check.py")
    assert res["n_gpus"] == 2 and res["ok"], res


@pytest.mark.parametrize("n,kernel", [(2048, "SquaredExpARD"), (1280, "MaternFiveHalves")])
def test_two_rank_distributed_fit_matches_lb_fit(n, kernel):
    """limbo_b200/dist_fit.py: the factor computed by two ranks together and assembled on every rank is the one lb_fit
    produces (bit-identical alpha, predictions, argmax; L compared element-wise for these sizes)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun(2, "tools/dist_fit_check.py", "--size", str(n), "--cands", "3000", "--kernel", kernel, "--reps", "1")
    assert res["n_gpus"] == 2 and res["supported"] and res["info"] == 0
    assert res["bit_identical_on_every_rank"], res
    assert res["loglik_rel_diff"] == 0.0, res


@pytest.mark.parametrize("precision", ["fp16", "fp16x3"])
def test_two_rank_distributed_inverse_matches_replicated(precision):
    """limbo_b200/dist_inv.py: each rank inverts its column tiles of the factor, one all_gather assembles the reduced-precision
    copy; variances agree with the replicated inversion and the sharded EI argmax is the unsharded one."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun(2, "tools/dist_inv_check.py", "--size", "4500", "--cands", "6000", "--dim", "6", "--precision", precision, "--reps", "1")
    assert res["n_gpus"] == 2 and res["supported"]
    assert res["ok_on_every_rank"], res
