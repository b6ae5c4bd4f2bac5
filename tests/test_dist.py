"""Multi-process tests of the row-block -> all-to-all -> column-block commit (lcpc_b200/dist.py).

CPU (not gpu): world_size 2 over gloo with a checker backend, covering the partition plan, the
uneven all-to-all split sizes, the aligned-subtree root assembly and the constant padding roots.
GPU: the same worker over NCCL with the CUDA backend (needs >= 2 GPUs; skipped otherwise).
"""
import json
import os
import socket
import subprocess
import sys

import pytest

from lcpc_b200.dist import make_plan

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_workers(world, backend, kind, field, n, seed=0, timeout=600, transport=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(HERE, "dist_worker.py"), backend, kind, str(field), str(n), str(seed)]
    if transport:
        cmd.append(transport)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    # ranks share one pipe: two records can end up on one line, so decode objects one after another
    lines, dec = [], json.JSONDecoder()
    for l in out.stdout.splitlines():
        pos = l.find("{")
        while pos >= 0:
            try:
                obj, end = dec.raw_decode(l, pos)
            except json.JSONDecodeError:
                break
            lines.append(obj)
            pos = l.find("{", end)
    assert out.returncode == 0 and len(lines) == world, out.stderr[-3000:]
    return sorted(lines, key=lambda d: d["rank"])


def check(lines):
    for d in lines:
        assert d["root"] == d["want"] == d["again"], d
        assert d["ok_cols"] and d["ok_leaves"], d
        assert d["ok_prove"] and d["ok_verify"], d  # prove() over the sharded commit == the oracle's proof


@pytest.mark.parametrize("shape", [(256, 65536, 131072, 8), (72, 235173, 357699, 8), (72, 235173, 357699, 4),
                                   (2, 512, 1024, 2), (64, 16384, 32768, 1), (286, 940690, 1430790, 8), (3, 8, 16, 8),
                                   (1, 300, 457, 2), (5, 1, 2, 4)])
def test_plan_partitions(shape):
    n_rows, n_per_row, n_cols, world = shape
    p = make_plan(n_rows, n_per_row, n_cols, world)
    assert p.row_lo[0] == 0 and p.row_lo[-1] == n_rows and p.col_lo[0] == 0 and p.col_lo[-1] == n_cols
    assert all(a <= b for a, b in zip(p.row_lo, p.row_lo[1:])) and all(a <= b for a, b in zip(p.col_lo, p.col_lo[1:]))
    assert p.n_sub * p.sub_leaves == p.np2 and p.n_sub & (p.n_sub - 1) == 0
    assert (p.n_real_sub - 1) * p.sub_leaves < n_cols <= p.n_real_sub * p.sub_leaves
    # column blocks are unions of whole subtrees (except the ragged tail) and balanced to one subtree
    widths = [p.sub_lo[g + 1] - p.sub_lo[g] for g in range(world)]
    assert max(widths) - min(widths) <= 1
    for g in range(world):
        assert p.col_lo[g] == min(p.sub_lo[g] * p.sub_leaves, n_cols)
    rows = [p.row_lo[g + 1] - p.row_lo[g] for g in range(world)]
    assert max(rows) - min(rows) <= 1


def test_gloo_world2_ligero():
    check(run_workers(2, "gloo", "ligero", 1, 1 << 12))


def test_gloo_world2_brakedown_non_power_of_two_columns():
    check(run_workers(2, "gloo", "sdig", 1, 3000, seed=1))


def test_gloo_world3_ragged():
    # 3 ranks, row count not divisible, Ft127
    check(run_workers(3, "gloo", "ligero", 2, (1 << 11) - 5))


def test_gloo_world4_brakedown_ft191():
    # 4 ranks, 3-limb field, non-power-of-two column count: commit + prove over the sharded commit
    check(run_workers(4, "gloo", "sdig", 3, 2500, seed=2))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["auto", "nccl"])
def test_nccl_ligero_ft255(transport):
    """auto = peer-mapped stores from the encode's last pass when the GPUs allow it; nccl = fused pack + all-to-all"""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    lines = run_workers(min(n, 8), "nccl", "ligero", 4, 1 << 16, transport=transport)
    check(lines)
    assert len({d["transport"] for d in lines}) == 1
    if transport == "nccl":
        assert lines[0]["transport"] == "nccl"


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["auto", "nccl"])
def test_nccl_brakedown_ft127(transport):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    check(run_workers(min(n, 8), "nccl", "sdig", 2, 1 << 15, seed=0, transport=transport))


@pytest.mark.gpu
def test_nccl_ligero_ft255_2_20_multi_pass():
    """2^15-point rows: the scatter store sits in the second NTT pass"""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    check(run_workers(min(n, 8), "nccl", "ligero", 4, 1 << 20))
