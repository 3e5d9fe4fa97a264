"""Drop-in contract of the nn.Modules (CPU): names, attributes, state_dict layout, default init."""
import importlib

import pytest
import torch

from helpers import spec, synth
from oracle import refshim

sm = importlib.import_module("3dal_pytorch_b200.static_model")
dm = importlib.import_module("3dal_pytorch_b200.dynamic_model")

CASES = [("static_one", sm.StaticModelOneBoxEst, 109), ("static_two", sm.StaticModelTwoBoxEst, 153),
         ("dynamic", dm.DynamicModel, 165)]


@pytest.mark.parametrize("kind,cls,n_tensors", CASES)
def test_state_dict_layout_follows_spec(kind, cls, n_tensors):
    m = cls()
    sd = m.state_dict()
    assert len(sd) == n_tensors                                   # 109 tensors for static-one
    want = synth.random_state_dict(kind, seed=0)
    assert set(sd) == set(want)
    for k in want:
        assert sd[k].shape == want[k].shape and sd[k].dtype == want[k].dtype, k
    assert m.load_state_dict(want).missing_keys == []


def test_attributes():
    a, b, c = sm.StaticModelOneBoxEst(), sm.StaticModelTwoBoxEst(), dm.DynamicModel()
    assert (a.name, b.name) == ("one_box_est", "two_box_est")
    assert (a.n_classes, a.n_channel, c.n_channel, c.r, c.s) == (3, 3, 4, 2, 50)
    assert sm.NUM_OBJECT_POINT == 512 and sm.NUM_POINT == 4096 and dm.NUM_POINT == 1024 and dm.NUM_FRAME == 5
    with pytest.raises(RuntimeError, match="no CPU path"):        # training mode is CUDA-only as well
        a.train()(torch.zeros(1, 3, 8), torch.zeros(1, 7), torch.zeros(1, 7))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("kind,cls,n_tensors", CASES)
def test_same_keys_and_same_default_init_as_reference(kind, cls, n_tensors):
    rs, rd, _, _ = refshim.load()
    ref_cls = {"static_one": rs.StaticModelOneBoxEst, "static_two": rs.StaticModelTwoBoxEst, "dynamic": rd.DynamicModel}[kind]
    torch.manual_seed(5)
    mine = cls().state_dict()
    torch.manual_seed(5)
    ref = ref_cls().state_dict()
    assert list(mine) != [] and set(mine) == set(ref)
    for k in ref:
        assert torch.equal(mine[k], ref[k]), k                    # same registration order => same RNG draws
    assert ref_cls().load_state_dict(mine).missing_keys == []


def test_fold_block_matches_eval_batchnorm():
    eng = importlib.import_module("3dal_pytorch_b200.engine")
    m = sm.StaticModelOneBoxEst()
    m.load_state_dict(synth.random_state_dict("static_one", seed=3))
    fw = eng.fold_block(m.box_est, m.box_est._table)
    x = torch.randn(4, 512)
    ref = torch.nn.functional.relu(m.box_est.fcbn1.eval()(m.box_est.fc1(x)))
    got = torch.relu(x @ fw["fc1"][0].t() + fw["fc1"][1])
    assert torch.allclose(ref, got, atol=1e-5)
