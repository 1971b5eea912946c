"""The closed form behind the compile-time kernels' SHIP scaling (csrc/scg_sc_static.cuh, ship_cuts_static), restated in
numpy and checked on the CPU against the oracle's sequential restatement of SC_Action.apply (supplychain_env.py:58-96).

The kernels never sort: with before(j, i) := a_j < a_i or (a_j == a_i and j < i) on the RAW fp32 actions,
    amount_i = (min(a_i, 1) - pred_i) * limit / 2,   pred_i = max(min(a_j, 1) for j before i), or -1 when nobody is,
which relies on two facts this file pins: the reference's "never more than what is left" clamp equals capping every cut
point at 1 (also with several actions above 1), and only the VALUE of the predecessor matters.  "Nobody before i" has to
stay a separate case: the start cut (raw -1) precedes every destination even when an out-of-range action lies below it.
The device code itself is exercised by the GPU parity tests; this is the algebra, runnable without a GPU."""
import numpy as np
import pytest

from oracle.sc_oracle import SupplyChainOracle


def _reference(raw, limit):
    values = [(float(a) + 1.0) / 2.0 for a in raw]                  # _denormalize_action, :697-698
    return np.array(SupplyChainOracle._ship_cuts(None, values, limit), dtype=np.float64)


def _before(raw, j, i):
    return raw[j] <= raw[i] if j < i else raw[j] < raw[i]


def _closed_form(raw, limit):
    """Any D: a select and a max per pair, -inf as "not before", then the separate "nobody" case."""
    raw = np.asarray(raw, dtype=np.float32)
    cpd = np.minimum(raw, np.float32(1))
    out = np.empty(len(raw))
    for i in range(len(raw)):
        m = np.float32(-np.inf)
        for j in range(len(raw)):
            if j != i:
                m = max(m, cpd[j] if _before(raw, j, i) else np.float32(-np.inf))
        pred = np.float32(-1) if m == -np.inf else m
        out[i] = (np.float64(cpd[i]) - np.float64(pred)) * (limit * 0.5)
    return out


def _closed_form_3(raw, limit):
    """D == 3 as the kernel writes it: three selects and one max per destination, no predicate logic."""
    raw = np.asarray(raw, dtype=np.float32)
    cpd = np.minimum(raw, np.float32(1))
    c01, c02, c12 = raw[0] <= raw[1], raw[0] <= raw[2], raw[1] <= raw[2]

    def prev_of(bj, bk, wj, wk):
        both = max(wj, wk)
        return (both if bk else wj) if bj else (wk if bk else np.float32(-1))

    pr = [prev_of(not c01, not c02, cpd[1], cpd[2]), prev_of(c01, not c12, cpd[0], cpd[2]), prev_of(c02, c12, cpd[0], cpd[1])]
    return np.array([(np.float64(cpd[i]) - np.float64(pr[i])) * (limit * 0.5) for i in range(3)])


def _random_actions(rs, D, kind):
    if kind == "in_range":
        a = rs.uniform(-1, 1, D)
    elif kind == "wide":
        a = rs.uniform(-1.6, 1.6, D)
    elif kind == "ties":
        a = rs.choice([-1.25, -1.0, -0.5, 0.0, 0.25, 1.0, 1.5], D)
    else:                                                            # a grid that hits every boundary often
        a = rs.choice([-1.0, 1.0, 0.999, -0.999, 2.0, -2.0, 0.0], D)
    return a.astype(np.float32)


@pytest.mark.parametrize("kind", ["in_range", "wide", "ties", "edges"])
@pytest.mark.parametrize("D", [1, 2, 3, 4, 5])
def test_closed_form_matches_the_sequential_reference(D, kind):
    rs = np.random.RandomState(100 * D + len(kind))
    for _ in range(400):
        raw = _random_actions(rs, D, kind)
        limit = float(rs.choice([1.0, 37.0, 812.5, 3000.0]))
        ref, got = _reference(raw, limit), _closed_form(raw, limit)
        assert np.allclose(got, ref, rtol=0, atol=4e-13 * limit), (raw, limit, got, ref)
        if D == 3:
            assert np.array_equal(_closed_form_3(raw, limit), got), raw


def test_every_ordering_and_tie_pattern_of_three():
    grid = np.array([-1.5, -1.0, -0.25, 0.5, 1.0, 1.75], dtype=np.float32)
    for a in grid:
        for b in grid:
            for c in grid:
                raw = np.array([a, b, c], dtype=np.float32)
                ref = _reference(raw, 100.0)
                assert np.allclose(_closed_form_3(raw, 100.0), ref, rtol=0, atol=1e-10), raw
                assert np.allclose(_closed_form(raw, 100.0), ref, rtol=0, atol=1e-10), raw


def test_start_cut_is_not_an_ordinary_predecessor():
    """max(candidates, -1) alone is wrong below -1: destination 0 at raw -1.5 precedes destination 1, whose amount is
    measured from -1.5 (not from the start cut -1), while destination 0's own (negative) amount is measured from -1."""
    raw = np.array([-1.5, 0.2, 0.5], dtype=np.float32)
    ref = _reference(raw, 10.0)
    assert ref[0] < 0 and np.isclose(ref[1], (0.2 + 1.5) * 5.0)
    assert np.allclose(_closed_form_3(raw, 10.0), ref, rtol=0, atol=1e-11)
    cpd = np.minimum(raw, np.float32(1))
    wrong_pred_1 = max(cpd[0], np.float32(-1))                       # the tempting two-select + one-max form
    assert not np.isclose((np.float64(cpd[1]) - np.float64(wrong_pred_1)) * 5.0, ref[1])
