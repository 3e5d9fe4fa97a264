"""CUDA-graph replay of the eval forward (3dal_pytorch_b200/graphs.py): bit-identical to the eager call."""
import importlib

import pytest
import torch

from helpers import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
sm = importlib.import_module("3dal_pytorch_b200.static_model")
dm = importlib.import_module("3dal_pytorch_b200.dynamic_model")
graphs = importlib.import_module("3dal_pytorch_b200.graphs")


def _inputs(kind, bs, seed):
    if kind == "dynamic":
        d = synth.dynamic_tracks(bs, seed=seed)
        return (torch.from_numpy(d["pts_pm"]).to(DEV).transpose(2, 1), torch.from_numpy(d["box_sm"]).to(DEV).transpose(2, 1), None)
    d = synth.static_tracks(bs, seed=seed)
    return (torch.from_numpy(d["pts_pm"]).to(DEV).transpose(2, 1), torch.from_numpy(d["init_box"]).to(DEV),
            torch.from_numpy(d["bbox_gt"]).to(DEV))


@pytest.mark.parametrize("kind,cls,bs", [("static_one", sm.StaticModelOneBoxEst, 32), ("static_two", sm.StaticModelTwoBoxEst, 32),
                                         ("dynamic", dm.DynamicModel, 16)])
@pytest.mark.parametrize("precision", ["mixed", "bf16x3", "fp32"])
def test_graph_replay_equals_eager(kind, cls, bs, precision):
    model = cls().to(DEV).eval()
    model.load_state_dict(synth.random_state_dict(kind, seed=11))
    model.precision = precision
    g = graphs.GraphedForward(model, *_inputs(kind, bs, seed=1))
    for seed in (2, 3):                                   # new data through the captured buffers
        inp = _inputs(kind, bs, seed)
        with torch.no_grad():
            eager = {k: v.clone() for k, v in model(*inp).items() if torch.is_tensor(v)}
        out = g(*inp)
        torch.cuda.synchronize()
        assert set(eager) <= set(out)
        for k, v in eager.items():
            assert torch.equal(out[k], v), (k, seed)


def test_graph_rejects_training_mode_and_shape_changes():
    model = sm.StaticModelOneBoxEst().to(DEV)
    model.load_state_dict(synth.random_state_dict("static_one", seed=11))
    with pytest.raises(ValueError):
        graphs.GraphedForward(model.train(), *_inputs("static_one", 4, 1))
    g = graphs.GraphedForward(model.eval(), *_inputs("static_one", 4, 1))
    with pytest.raises(ValueError):
        g(*_inputs("static_one", 8, 1))
