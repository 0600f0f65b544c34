"""The experimental two-tiles-per-CTA attention kernel (csrc/attention_tc2.cu, opt-in) has not run on a GPU yet; its
mbarrier protocol is model-checked on the CPU instead (tools/sim_attention_v2.py: coroutine model of the producer /
MMA / softmax roles with parity semantics, asynchronous TMA and in-order tensor-core completions, adversarial random
scheduling, hazard checks on every shared resource)."""
import os
import sys

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_protocol_model_has_no_hazard_or_deadlock():
    import random
    import sim_attention_v2 as sim
    for seed in range(250):
        rng = random.Random(10_000 + seed)
        window = rng.choice([-1, 64, 64, 0, 200])
        sim.Sim(sim.random_items(rng, window), seed).run()


def test_protocol_model_detects_a_missing_wait():
    """The model is sensitive: without the softmax warps' arrivals on q_empty (barrier count 1, only the MMA warp's
    commit arrives) a lagging warp of an absent tile aliases the parity of q_full -- the reason those arrivals exist in
    the kernel.  Some schedule among 120 must expose it."""
    import random
    import sim_attention_v2 as sim

    failures = 0
    for seed in range(120):
        rng = random.Random(10_000 + seed)
        s = sim.Sim(sim.random_items(rng, rng.choice([-1, 64, 0])), seed)
        s.q_empty = sim.Bar("q_empty", 1)
        commit = s.q_empty.arrive

        def arrive(tx=0, s=s, commit=commit):
            if s._running == "tc":        # the MMA warp's tcgen05.commit (runs from the tensor-core queue)
                commit(tx)                # ... arrivals issued by a softmax coroutine are dropped
        s.q_empty.arrive = arrive
        try:
            s.run()
        except AssertionError:
            failures += 1
    assert failures > 0
