"""The synchronisation protocol of the fused MLP kernel (motionbert_b200/csrc/mlp_fused.cuh) under a randomised scheduler:
tests/mlp_pipeline_model.py restates one CTA pair's producer / MMA / epilogue warps over modelled mbarriers, an in-order
tensor pipe and asynchronous TMA loads / stores, and asserts the data-flow properties (RAW / WAR on the hidden ring,
stage and accumulator hand-over, staging reuse, every output chunk once, no deadlock).  CPU only; the arithmetic itself is
checked bit for bit on the GPU (tests/test_gpu_mlp_fused.py)."""
import pytest

from mlp_pipeline_model import Sim

GEOMETRIES = {                       # (NT1 = hidden / 256, NT2 = C / 256, KB1 = C / 32 shortened: the protocol does not depend on it)
    "base (C=512, hidden=1024)": (4, 2, 4),
    "lite (C=256, hidden=1024)": (4, 1, 2),
    "wide (C=512, hidden=2048)": (8, 2, 3),
    "square (C=256, hidden=256)": (1, 1, 2),
    "C=768, hidden=768": (3, 3, 5),
}


@pytest.mark.parametrize("ring", [True, False])
@pytest.mark.parametrize("rounds", [1, 2, 3])
@pytest.mark.parametrize("geom", list(GEOMETRIES))
def test_protocol_holds_under_random_interleavings(geom, rounds, ring):
    NT1, NT2, KB1 = GEOMETRIES[geom]
    for seed in range(6):
        Sim(NT1, NT2, KB1, rounds, seed, ring=ring).run()


@pytest.mark.parametrize("bug,needle", [("no_hready_wait", "RAW"), ("early_signal", "RAW"), ("no_wait_read", "staging")])
def test_model_catches_broken_protocols(bug, needle):
    """The model has teeth: each deliberately broken variant trips the matching assertion.  (With short store latencies the
    pipeline usually satisfies the hidden-ready dependency by itself -- ncu: the producer waits for the flag in 5 of 29865
    samples -- so these runs draw the store latencies from a long tail; the correct protocol passes under the same tail.)"""
    kw = dict(ring=True, store_read=(1, 400), store_write=(3, 4000))
    caught = 0
    for seed in range(10):
        Sim(4, 2, 4, 2, seed, **kw).run()                        # the kernel's protocol: fine under the long tail too
        try:
            Sim(4, 2, 4, 2, seed, bug=bug, **kw).run()
        except AssertionError as e:
            assert needle in str(e), str(e)
            caught += 1
    assert caught >= 3, f"only {caught} of 10 interleavings exposed '{bug}'"
