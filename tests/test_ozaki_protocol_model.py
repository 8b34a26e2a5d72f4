"""Exhaustive exploration (every interleaving of a small instance) of the barrier protocol of the int8 update kernel -- producer, two
MMA issuers on alternate stages, epilogue, asynchronous copy / commit completions -- through tools/ozaki_protocol_model.py: no deadlock,
no stage consumed a phase early or overwritten under running MMAs, accumulators initialised first, epilogues only after all MMAs of a
pass.  The two-issuer variant on an ODD stage count fails (as it did on the GPU), which is why the kernel falls back to one issuer there."""
import pytest

from tools.ozaki_protocol_model import Violation, explore


@pytest.mark.parametrize("nq,nst_a,nst_b", [(8, 6, 4), (16, 6, 4), (12, 6, 3), (8, 4, 2), (8, 6, 6)])
def test_protocol_is_safe_and_live(nq, nst_a, nst_b):
    res = explore(nq, nst_a, nst_b)
    assert res["final_states"] == 1 and res["states"] > 1000


def test_two_issuers_on_an_odd_ring_are_caught():
    with pytest.raises(Violation, match="consumes stage"):
        explore(8, 6, 3, force_dual=True)
