"""The drop-in against the REAL reference class (build container only: needs the read-only checkout at /root/reference).

* ``predict`` / ``get_mask_proposals`` keep the reference's parameter names, order and defaults
  (networks/zutis.py:340-353, :177-182); ``precision`` is the one additive keyword.
* ``install(ZUTIS)`` leaves everything ``forward`` calls alone, and ``install(ZUTIS, inference_forward_ops=True)``
  routes autograd-recording calls to the reference's own methods: ``forward`` followed by ``backward`` still yields
  gradients for the parameters behind both criterion inputs (trainer.py:136-150).
"""
import importlib.util
import inspect
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "networks")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ZUTIS():
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.install_stubs()                       # clip / pycocotools stubs of SURVEY A.6, puts /root/reference on sys.path
    from networks.zutis import ZUTIS as cls
    saved = {k: cls.__dict__[k] for k in ("predict", "get_mask_proposals", "image_to_text_space")}
    yield cls
    for k, v in saved.items():
        setattr(cls, k, v)
    if REF in sys.path:
        sys.path.remove(REF)


def _params(fn, drop=()):
    return [(p.name, p.default, p.kind) for p in inspect.signature(fn).parameters.values() if p.name not in drop]


def test_signatures_match_the_reference(ZUTIS):
    from zutis_b200 import decode
    ref_predict = inspect.unwrap(ZUTIS.predict)                      # the reference wraps it in @torch.no_grad()
    assert _params(inspect.unwrap(decode.predict), drop=("precision",)) == _params(ref_predict)
    assert _params(decode.get_mask_proposals, drop=("precision",)) == _params(ZUTIS.get_mask_proposals)
    assert _params(decode.image_to_text_space, drop=("precision",)) == _params(ZUTIS.image_to_text_space)
    from zutis_b200 import RunningScore, compute_iou
    sys_path_had = REF in sys.path
    from utils.running_score import RunningScore as RefScore
    from utils.iou import compute_iou as ref_iou
    assert sys_path_had
    assert _params(RunningScore.update) == _params(RefScore.update)
    assert _params(RunningScore.get_scores) == _params(RefScore.get_scores)
    assert _params(RunningScore.reset) == _params(RefScore.reset)
    assert _params(RunningScore.__init__, drop=("device",)) == _params(RefScore.__init__)
    assert _params(compute_iou) == _params(ref_iou)


@pytest.mark.parametrize("forward_ops", [False, True])
def test_training_keeps_its_gradients_after_install(ZUTIS, forward_ops):
    import zutis_b200
    originals = (ZUTIS.get_mask_proposals, ZUTIS.image_to_text_space)
    zutis_b200.install(ZUTIS, inference_forward_ops=forward_ops)
    if not forward_ops:
        assert (ZUTIS.get_mask_proposals, ZUTIS.image_to_text_space) == originals
    cats = ["background"] + [f"category number {i}" for i in range(4)]
    net = ZUTIS(categories=cats, clip_arch="ViT-B/32", device=torch.device("cpu")).train()
    torch.manual_seed(0)
    out = net(torch.randn(1, 3, 64, 64))
    assert out["mask_proposals"].requires_grad and out["patch_tokens"].requires_grad
    (out["mask_proposals"].mean() + out["patch_tokens"].square().mean()).backward()
    with_grad = [n for n, p in net.named_parameters() if p.grad is not None and p.grad.abs().sum() > 0]
    assert any(n.startswith("ffn2") or "decoder" in n for n in with_grad), "mask-proposal branch lost its graph"
    assert any("encoder" in n or "proj" in n for n in with_grad), "patch-token branch lost its graph"
