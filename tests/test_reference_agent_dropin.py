"""CPU (build container only -- skipped where /root/reference is absent): the reference's own agent code
(agents/decentralplannerlocal.py, utils/multirobotsim_dcenlocal.py, utils/metrics.py, graphs/losses/*),
imported UNMODIFIED after `install_dropin()`, must import and run against this package's modules.

The product has no CPU path, so for this CPU-box test the three native entry points of the drop-in
module are replaced by the oracle (test-only monkeypatch): what is under test here is the boundary
-- import routing, constructor / .to() / addGSO / forward / list-of-N return / train() / eval() /
parameters() / optimizer interplay exactly as the reference's agent drives them -- and the result must
reproduce the golden trace the same agent code produced with the reference's own model
(tests/golden/make_agent_trace.py).  The CUDA numerics of the same calls are pinned by
tests/test_gpu_agent_trace.py against the same trace.
"""
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_agent
from oracle import planner_oracle as po

pytestmark = pytest.mark.skipif(not ref_agent.available(), reason="reference tree not present (GPU box)")


def test_install_dropin_keeps_the_reference_packages_importable():
    import gnn_pathplanning_b200 as gp
    with ref_agent.reference_env(dropin=True) as agmod:
        # the three routed names
        import graphs.models.decentralplanner as dcp
        import graphs.weights_initializer as wi
        import utils.graphUtils.graphML as gml
        assert dcp.DecentralPlannerNet is gp.DecentralPlannerNet and agmod.DecentralPlannerNet is gp.DecentralPlannerNet
        assert gml.GraphFilterBatch is gp.GraphFilterBatch and gml.BatchLSIGF is gp.BatchLSIGF
        assert wi.weights_init is gp.weights_init
        # everything else still comes from the reference tree (the round-1 shim shadowed these)
        import graphs.losses.cross_entropy as ce
        import graphs.losses.regularizer  # noqa: F401
        import utils.metrics as metrics
        import utils.misc  # noqa: F401
        import utils.multirobotsim_dcenlocal as sim
        import utils.graphUtils.graphTools as gt
        for m in (ce, metrics, sim, gt, agmod):
            assert m.__file__.startswith(ref_agent.REF_ROOT), m.__file__
        assert agmod.CrossEntropyLoss is ce.CrossEntropyLoss
        assert agmod.multiRobotSim is sim.multiRobotSim
        # the OnlineExpert agent (same model calls) is imported by agents/__init__.py as well
        assert "agents.decentralplannerlocal_OnlineExpert" in sys.modules
    assert "graphs.models.decentralplanner" not in sys.modules          # uninstall() cleaned up


def test_install_dropin_replaces_already_imported_reference_modules():
    import gnn_pathplanning_b200 as gp
    with ref_agent.reference_env(dropin=False):
        import graphs.models.decentralplanner as ref_dcp
        assert ref_dcp.DecentralPlannerNet is not gp.DecentralPlannerNet
        gp.install_dropin()
        try:
            import graphs.models.decentralplanner as dcp
            import graphs.models
            assert dcp.DecentralPlannerNet is gp.DecentralPlannerNet
            assert graphs.models.decentralplanner is dcp
        finally:
            from gnn_pathplanning_b200 import dropin
            dropin.uninstall()


def _oracle_backed(monkeypatch):
    """Test-only: route the drop-in module's native calls to the oracle so the boundary can be driven on CPU."""
    from gnn_pathplanning_b200 import planner as pl, graphml

    monkeypatch.setattr(pl, "_require_cuda", lambda t, name: None)
    monkeypatch.setattr(graphml, "_require_cuda", lambda t, name: None)

    def fused(self, x, S):
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        return torch.stack(po.planner_forward(sd, S[:, 0], x.float()))

    class TrainFn:
        @staticmethod
        def apply(module, x, S, *params):
            sd = dict(module.named_parameters())
            bn = {k: v for k, v in module.named_buffers()}
            full = dict(sd)
            full.update(bn)
            return torch.stack(po.planner_forward(full, S[:, 0], x.float(), True, bn))

    monkeypatch.setattr(pl.DecentralPlannerNet, "_forward_fused", fused)
    monkeypatch.setattr(pl, "_PlannerTrainFn", TrainFn)


def test_reference_agent_runs_unchanged_against_the_dropin(golden, monkeypatch):
    g = golden("agent_trace.npz")
    N, K = int(g["N"]), int(g["K"])
    sd = {k[3:]: torch.from_numpy(np.ascontiguousarray(g[k])) for k in g.files if k.startswith("sd_")}
    batch = (torch.from_numpy(g["batch_x"]), torch.from_numpy(g["batch_target"]), torch.zeros(int(g["B"])),
             torch.from_numpy(g["batch_S"]), torch.zeros(int(g["B"]), 1))
    case = tuple(torch.from_numpy(g[k]) for k in ("case_input", "case_target", "case_makespan", "case_map"))
    cfg = ref_agent.make_config(N, K, "cpu")
    _oracle_backed(monkeypatch)
    with ref_agent.reference_env(dropin=True) as agmod:
        import gnn_pathplanning_b200 as gp
        model = agmod.DecentralPlannerNet(cfg)                  # agents/decentralplannerlocal.py:47
        assert isinstance(model, gp.DecentralPlannerNet)
        model.load_state_dict(sd)
        model = model.to(cfg.device)                            # :88
        tr = ref_agent.run_agent_trace(agmod, model, cfg, batch, case)
    assert abs(float(tr["train_loss"]) - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    assert rel_err(tr["train_logits"], g["train_logits"]) <= 1e-5
    assert np.array_equal(tr["rollout_actions"], g["rollout_actions"])
    assert np.array_equal(tr["rollout_pos"], g["rollout_pos"])
    assert np.array_equal(tr["rollout_x"], g["rollout_x"]) and np.array_equal(tr["rollout_S"], g["rollout_S"])
    assert rel_err(tr["rollout_logits"], g["rollout_logits"]) <= 1e-5
