"""CPU, build container only: oracle/sim_oracle.py against the reference's own multiRobotSim (unmodified), step by
step on seeded cases -- FOV tensors, float64 GSOs (incl. the step-0 radius growth), positions, actions, flags, goal /
step bookkeeping.  `random.choice` is the round-robin contract on both sides (oracle/ref_sim.py)."""
import numpy as np
import pytest

from oracle import ref_agent, ref_sim

pytestmark = pytest.mark.skipif(not ref_agent.available(), reason="reference tree not present (GPU box)")

CASES = [  # N, W, obstacle density, makespan (-> maxstep = 2x, 3x for N >= 20), steps, seed
    (10, 20, 0.10, 14, 28, 1), (10, 20, 0.10, 6, 12, 2), (5, 8, 0.25, 10, 20, 3), (6, 6, 0.05, 8, 16, 4),
    (20, 28, 0.10, 6, 18, 5), (3, 12, 0.3, 9, 18, 6), (12, 9, 0.0, 5, 10, 7),
]


@pytest.mark.parametrize("N,W,density,makespan,T,seed", CASES)
def test_oracle_matches_reference_simulator(N, W, density, makespan, T, seed):
    rng = np.random.default_rng(1000 + seed)
    case = ref_sim.make_case(rng, N, W, density, makespan)
    with ref_agent.reference_env(dropin=False) as agmod:
        ref = ref_sim.reference_rollout(agmod, case, T, seed, N)
    got = ref_sim.oracle_rollout(case, T, seed, N)
    for k in ref:
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k
    assert (ref["last_action"] != 4).any()                       # agents really moved
