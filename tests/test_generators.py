"""Structure of the generated task lists (algebraist_b200/csrc/gen_col.py, gen_tz.py): every column / coset / parent is
owned by exactly one task, and the task-size caps the kernels were tuned with hold.  (What the generated code COMPUTES
is checked against the oracle by the emulation and GPU suites.)"""
import os
import sys

import pytest

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "algebraist_b200", "csrc")
sys.path.insert(0, CSRC)
import gen_col  # noqa: E402
import gen_tz  # noqa: E402
from gen_base import children, coef_off, dim, partitions  # noqa: E402


@pytest.mark.parametrize("inverse", [False, True])
def test_s6_items_cover_every_level5_column_once(inverse):
    _, items = gen_col.gen_s6(inverse)
    mus = list(partitions(5))
    seen = sorted((task, c) for _, task, c, _ in items)
    assert seen == sorted((t, c) for t, mu in enumerate(mus) for c in range(dim(mu)))
    assert len(items) == 26
    for _, task, c, off in items:
        assert off == coef_off(mus[task]) + c
    assert [it[0] for it in items] == sorted((it[0] for it in items), reverse=True)      # most expensive first


@pytest.mark.parametrize("k", [7, 8])
def test_inverse_coset_runs_partition_the_cosets_and_respect_the_sums_cap(k):
    for mu in partitions(k - 1):
        size = gen_col.inv_cosets(k, dim(mu))
        assert 1 <= size <= gen_col.INV_COSETS_PER_TASK[k]
        assert size == 1 or size * dim(mu) <= gen_col.INV_SUMS_CAP[k]
        runs = gen_col.coset_groups(k, size)
        assert [i for r in runs for i in r] == list(range(k))


def test_forward_parent_groups_cover_every_parent_once():
    cap = gen_col.FWD_ROWS_CAP[7]
    for mu in partitions(6):
        groups = gen_col.parent_groups(mu, cap)
        flat = [lam for g in groups for lam in g]
        assert sorted(flat) == sorted(lam for lam in partitions(7) if mu in children(lam))
        for g in groups:
            assert len(g) == 1 or sum(dim(lam) for lam in g) <= cap


@pytest.mark.parametrize("k", [9, 10, 11])
def test_t_tasks_cover_every_output_row_once(k):
    """The T tasks (mu, alpha) of level k own, over all parents lambda of mu and all paths lambda -> alpha, every row of
    the block mu of every lambda exactly once."""
    MB = gen_tz.LEVEL_MB[k]
    s = k - MB
    tasks = set(gen_tz.t_tasks(k))
    for lam in partitions(k):
        rows = 0
        for Pp, offp in gen_tz.subblocks(lam, s):
            rows += dim(Pp[-1])
        assert rows == dim(lam)
    for mu in partitions(k - 1):
        alphas = {alpha for (m, alpha) in tasks if m == mu}
        want = set()
        for lam in gen_tz.parents(mu):
            want |= {Pp[-1] for Pp, _ in gen_tz.subblocks(lam, s)}
        assert alphas == want
