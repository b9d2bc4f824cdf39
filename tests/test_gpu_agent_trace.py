"""GPU: the calls the reference's own agent made (recorded by tests/golden/make_agent_trace.py from
agents/decentralplannerlocal.py train_one_epoch :276-326 and mutliAgent_ActionPolicy :535-648 running against the
reference's model) replayed against the CUDA module: same call sequence, same dtypes (float64 [1,N,N] GSO per
rollout step), logits within 1e-5, the simulator's per-agent action choice (LogSoftmax + argmax,
utils/multirobotsim_dcenlocal.py:589-591) identical, training loss identical."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


class Cfg:
    def __init__(self, n, k):
        self.num_agents, self.nGraphFilterTaps = n, k


def _load(golden, prefix="sd_"):
    import gnn_pathplanning_b200 as gp
    g = golden("agent_trace.npz")
    N, K = int(g["N"]), int(g["K"])
    sd = {k[len(prefix):]: torch.from_numpy(np.ascontiguousarray(g[k])) for k in g.files if k.startswith(prefix)}
    cfg = Cfg(N, K)
    model = gp.DecentralPlannerNet(cfg)             # constructed before config.device exists (:47 vs :86)
    model.load_state_dict(sd)
    cfg.device = torch.device("cuda")
    return g, cfg, model.to(cfg.device)


def test_train_one_epoch_batch(golden):
    g, cfg, model = _load(golden)
    loss_fn = torch.nn.CrossEntropyLoss().to(cfg.device)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    model.train()
    inputGPU = torch.from_numpy(g["batch_x"]).to(cfg.device)
    gsoGPU = torch.from_numpy(g["batch_S"]).to(cfg.device)
    batch_targetGPU = torch.from_numpy(g["batch_target"]).to(cfg.device).permute(1, 0, 2)
    opt.zero_grad()
    loss = 0
    model.addGSO(gsoGPU)
    predict = model(inputGPU)
    for id_agent in range(cfg.num_agents):
        loss = loss + loss_fn(predict[id_agent][:], torch.max(batch_targetGPU[id_agent][:][:], 1)[1])
    loss = loss / cfg.num_agents
    loss.backward()
    opt.step()
    assert rel_err(torch.stack(predict).detach().cpu().numpy(), g["train_logits"]) <= 1e-5
    assert abs(loss.item() - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    # BatchNorm running statistics after the step (the parameters themselves took one Adam step: sign(g) * lr,
    # which amplifies rounding noise on near-zero gradients and is therefore not compared element-wise)
    after = model.state_dict()
    for k in after:
        if "running" in k:
            assert rel_err(after[k].double().cpu().numpy(), g["after_" + k]) <= 1e-5, k


@pytest.mark.parametrize("gf_mode", ["auto", "tc"])
def test_rollout_steps(golden, gf_mode):
    g, cfg, model = _load(golden, "after_")        # the agent rolled out with the weights its training step left
    model.eval()
    model.set_graph_filter_mode(gf_mode)
    logsm = torch.nn.LogSoftmax(dim=-1)
    T = g["rollout_x"].shape[0]
    worst = 0.0
    with torch.no_grad():
        for t in range(T):
            currentStateGPU = torch.from_numpy(g["rollout_x"][t]).to(cfg.device)       # [1,N,3,11,11]
            gso = torch.from_numpy(g["rollout_S"][t])                                  # float64 [1,N,N]
            assert gso.dtype == torch.float64
            model.addGSO(gso.to(cfg.device))
            actionVec_predict = model(currentStateGPU)
            assert isinstance(actionVec_predict, list) and len(actionVec_predict) == cfg.num_agents
            got = torch.stack(actionVec_predict).cpu().numpy()
            worst = max(worst, rel_err(got, g["rollout_logits"][t]))
            keys = [int(torch.max(logsm(actionVec_predict[i]), 1)[1]) for i in range(cfg.num_agents)]
            ref_keys = g["rollout_logits"][t][:, 0].argmax(-1)
            assert keys == ref_keys.tolist(), t
    assert worst <= 1e-5
