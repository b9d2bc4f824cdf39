"""GPU: the device-resident rollout step (csrc/rollout.cu, rows f1 / f2) against rollouts of the reference's own
simulator (tests/golden/rollout_trace.npz) and against oracle/sim_oracle.py -- bit-exact: FOV tensors, float64 GSOs incl.
the step-0 radius growth, moves with edge / obstacle / vertex-conflict / swap shielding, goal and step bookkeeping."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_inputs_vs_reference_builders(golden):
    """AgentState.toInputTensor and computeAdjacencyMatrix outputs of the reference (tests/golden/inputs.npz)."""
    import gnn_pathplanning_b200 as gp
    g = golden("inputs.npz")
    for i in range(int(g["n"])):
        starts, goals, m = g["starts%d" % i], g["goals%d" % i], g["map%d" % i]
        ro = gp.BatchedRollout(starts.shape[0], 6.0).setup(starts[None], goals[None], m, 10)
        x, S = ro.build_inputs(1)                      # step >= 1: fixed radius (the golden's branch)
        assert np.array_equal(x[0].cpu().numpy().astype(np.uint8), g["fov%d" % i])
        assert np.array_equal(S[0].cpu().numpy(), g["gso%d" % i])      # float64, bit for bit


def test_replay_of_reference_rollouts(golden):
    import gnn_pathplanning_b200 as gp
    g = golden("rollout_trace.npz")
    for name in g["sets"]:
        name = str(name)
        N, W, makespan, T, E = [int(v) for v in g[name + "_cfg"]]
        ro = gp.BatchedRollout(N, 6.0).setup(g[name + "_start"], g[name + "_goal"], g[name + "_map"], g[name + "_maxstep"])
        for t in range(T):
            x, S = ro.build_inputs(t)
            assert np.array_equal(x.cpu().numpy().astype(np.uint8), g[name + "_x"][:, t]), (name, t)
            assert np.array_equal(S.cpu().numpy(), g[name + "_S"][:, t]), (name, t)
            assert np.array_equal(ro.radius.cpu().numpy(), g[name + "_radius"][:, t]), (name, t)
            logits = torch.from_numpy(np.ascontiguousarray(g[name + "_logits"][:, t].transpose(1, 0, 2))).cuda()   # [N,E,5]
            flags = ro.move(logits, t + 1)
            assert np.array_equal(flags.cpu().numpy(), g[name + "_flags"][:, t]), (name, t)
            assert np.array_equal(ro.pos.cpu().numpy(), g[name + "_pos"][:, t]), (name, t)
            assert np.array_equal(ro.last_action.cpu().numpy(), g[name + "_last_action"][:, t]), (name, t)
            assert np.array_equal(ro.reached.cpu().numpy(), g[name + "_reached"][:, t]), (name, t)
        assert np.array_equal(ro.start_step.cpu().numpy(), g[name + "_start_step"])
        assert np.array_equal(ro.end_step.cpu().numpy(), g[name + "_end_step"])
        assert np.array_equal(ro.choice_counter.cpu().numpy(), g[name + "_choices"])


class Cfg:
    def __init__(self, n, k):
        self.num_agents, self.nGraphFilterTaps, self.device = n, k, torch.device("cuda")


@pytest.mark.parametrize("N,W,B,maxstep", [(10, 20, 64, 30), (6, 10, 33, 18), (32, 30, 5, 10), (40, 36, 3, 8)])
def test_run_loop_with_planner_vs_oracle(N, W, B, maxstep):
    """BatchedRollout.run: B episodes advanced on the device by the CUDA planner (no per-step host copies); the same
    episodes advanced one by one by oracle/sim_oracle.py fed with the same logits; identical trajectories, goal flags
    and step bookkeeping, incl. the freezing of finished episodes (agents/decentralplannerlocal.py:606-613).  N <= 32
    runs the warp-per-episode move kernel, N = 40 the thread-per-episode one."""
    import gnn_pathplanning_b200 as gp
    from gnn_pathplanning_b200 import synthetic
    from oracle import planner_oracle as po, sim_oracle
    rng = np.random.default_rng(N * W)
    cases = [synthetic.random_episode(rng, N, W, 0.1) for _ in range(B)]
    maps = np.stack([c[0] for c in cases]); starts = np.stack([c[1] for c in cases]); goals = np.stack([c[2] for c in cases])
    sd = po.init_state_dict(3, seed=3)
    po.randomize_bn_stats(sd, seed=4)
    sd["actionsMLP.0.weight"] = sd["actionsMLP.0.weight"] * 40.0          # a policy that moves (see make_agent_trace.py)
    m = gp.DecentralPlannerNet(Cfg(N, 3))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    ro = gp.BatchedRollout(N, 6.0).setup(starts, goals, maps, maxstep)
    sims = [sim_oracle.SimOracle(N, 6.0).setup(starts[b], goals[b], maps[b], maxstep) for b in range(B)]
    alive = [True] * B
    with torch.no_grad():
        for step in range(maxstep):
            x, S = ro.build_inputs(step)
            m.addGSO(S)
            logits = m.forward_logits(x)
            lg = logits.cpu().numpy()
            active_before = ro.active.cpu().numpy().copy()
            flags = ro.move(logits, step + 1)
            ro.active = (ro.active.bool() & ~flags[:, 0].bool() & (ro.maxstep > step + 1)).to(torch.int32)
            for b in range(B):
                assert bool(active_before[b]) == alive[b], (step, b)
                if not alive[b]:
                    continue
                xo, So = sims[b].inputs(step)
                assert np.array_equal(xo, x[b].cpu().numpy()) and np.array_equal(So, S[b].cpu().numpy()), (step, b)
                f = sims[b].move(lg[:, b], step + 1)
                if f[0] or step + 1 >= maxstep:
                    alive[b] = False
            pos = ro.pos.cpu().numpy()
            for b in range(B):
                assert np.array_equal(pos[b], np.array(sims[b].cur)), (step, b)
    assert np.array_equal(ro.reached.cpu().numpy(), np.array([[int(v) for v in s.reached] for s in sims]))
    assert np.array_equal(ro.end_step.cpu().numpy(), np.array([[-1 if v is None else v for v in s.end_step] for s in sims]))
    # the packaged loop gives the same final state
    ro2 = gp.BatchedRollout(N, 6.0).setup(starts, goals, maps, maxstep)
    steps = ro2.run(m, poll_every=4)
    assert steps <= maxstep and np.array_equal(ro2.pos.cpu().numpy(), ro.pos.cpu().numpy())
    assert np.array_equal(ro2.reached.cpu().numpy(), ro.reached.cpu().numpy())
