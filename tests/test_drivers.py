"""The drivers (exploring_meta_b200/vision/*.py) against the reference's outer loop restated with the oracle:
parameters after two meta-iterations (meta-gradient / B -> Adam), logged metrics, BatchNorm running statistics in the
reference's call order, checkpoint key names.  Runs on the CPU emulator of the C ABI and (-m gpu) on cuda:0."""
import os

import pytest
import torch

from exploring_meta_b200.synthetic import get_tasks
from oracle import maml_oracle as mo


@pytest.fixture
def in_tmp(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_maml_driver_two_iterations(kdev, in_tmp, monkeypatch):
    from exploring_meta_b200.vision import maml_vision as drv
    monkeypatch.setattr(drv, 'dataset', 'omni')
    monkeypatch.setattr(drv, 'cuda', kdev.type == 'cuda')
    P = dict(ways=5, shots=1, outer_lr=0.003, inner_lr=0.4, adapt_steps=1, meta_batch_size=3, num_iterations=2,
             save_every=1, seed=42)
    monkeypatch.setattr(drv, 'params', dict(P))
    exp = drv.MamlVision()
    # ---- the same two iterations with the oracle (vision/maml_vision.py:95-141) -------------------------------
    train, valid, test = get_tasks('omni', 5, 1, seed=42)
    ospec = mo.omniglot_spec(5)
    theta = [p.double() for p in mo.init_params(ospec, seed=42)]
    state = mo.new_adam_state(theta)
    rm = [torch.zeros(64, dtype=torch.float64) for _ in range(4)]
    rv = [torch.ones(64, dtype=torch.float64) for _ in range(4)]
    for it in range(2):
        xv, yv = valid.sample_batch(3)
        xt, yt = train.sample_batch(3)
        out_t = mo.meta_iteration(theta, xt.double(), yt, ospec, 1, 0.4)
        out_v = mo.meta_iteration(theta, xv.double(), yv, ospec, 1, 0.4)
        # BN calls: per task, the train task's calls then the validation task's calls (2 calls x 4 layers each)
        per_task = 2 * 4
        calls = []
        for t in range(3):
            calls += out_t['bn_calls'][t * per_task:(t + 1) * per_task] + out_v['bn_calls'][t * per_task:(t + 1) * per_task]
        rm, rv = mo.compose_running_stats(rm, rv, calls)
        assert exp.metrics['train_loss'][it] == pytest.approx(float(out_t['loss'].mean()), rel=1e-4)
        assert exp.metrics['valid_loss'][it] == pytest.approx(float(out_v['loss'].mean()), rel=1e-4)
        assert exp.metrics['train_acc'][it] == pytest.approx(float(out_t['correct'].sum()) / 15, abs=1e-6)
        assert exp.metrics['valid_acc'][it] == pytest.approx(float(out_v['correct'].sum()) / 15, abs=1e-6)
        theta = mo.adam_step(theta, [g / 3 for g in out_t['grad']], state, lr=0.003)
    got = mo.flatten([p.detach().cpu() for p in exp.model.parameters()])
    assert mo.rel_l2(got, mo.flatten(theta)) < 1e-5
    # the final evaluate() (vision/maml_vision.py:156) forwards meta_batch_size test tasks through the shared buffers too
    got_e = [test.sample() for _ in range(3)]                 # evaluate() calls test_tasks.sample() per task
    xe, ye = torch.stack([g[0] for g in got_e]), torch.stack([g[1] for g in got_e])
    out_e = mo.meta_iteration(theta, xe.double(), ye, ospec, 1, 0.4)
    rm, rv = mo.compose_running_stats(rm, rv, out_e['bn_calls'])
    assert exp.metrics['test_acc'][0] == pytest.approx(float(out_e['correct'].sum()) / 15, abs=1e-6)
    for l, blk in enumerate(exp.model.base):
        assert int(blk.normalize.num_batches_tracked) == (2 * 2 + 1) * 3 * 2
        assert torch.allclose(blk.normalize.running_mean.cpu().double(), rm[l], rtol=1e-4, atol=1e-5)
        assert torch.allclose(blk.normalize.running_var.cpu().double(), rv[l], rtol=1e-4, atol=1e-5)
    # checkpoints carry the reference's state_dict keys
    sd = torch.load(os.path.join(exp.model_path, 'model.pt'), map_location='cpu')
    assert 'base.0.normalize.running_mean' in sd and 'base.3.conv.weight' in sd and 'linear.bias' in sd
    assert os.path.isfile(os.path.join(exp.model_path, 'model_checkpoints', 'model_0.pt'))
    assert os.path.isfile(os.path.join(exp.model_path, 'metrics.json'))
    assert 'test_acc' in exp.metrics


def test_anil_driver_two_iterations(kdev, in_tmp, monkeypatch):
    from exploring_meta_b200.vision import anil_vision as drv
    monkeypatch.setattr(drv, 'dataset', 'omni')
    monkeypatch.setattr(drv, 'cuda', kdev.type == 'cuda')
    from exploring_meta_b200.vision import maml_vision
    monkeypatch.setattr(maml_vision, 'cuda', kdev.type == 'cuda')
    P = dict(ways=5, shots=1, outer_lr=0.003, inner_lr=0.5, adapt_steps=2, meta_batch_size=2, num_iterations=2,
             save_every=1000, seed=42)
    monkeypatch.setattr(drv, 'params', dict(P))
    exp = drv.AnilVision()
    train, valid, _test = get_tasks('omni', 5, 1, seed=42)
    ospec = mo.NetSpec(1, 28, 28, 32, 5, 4, False, 'flatten')
    body, head = mo.init_anil_params(ospec, seed=42)
    body, head = [p.double() for p in body], [p.double() for p in head]
    state = mo.new_adam_state(body + head)
    for it in range(2):
        xv, yv = valid.sample_batch(2)
        xt, yt = train.sample_batch(2)
        out = mo.meta_iteration(body, xt.double(), yt, ospec, 2, 0.5, anil_head=head)
        out_v = mo.meta_iteration(body, xv.double(), yv, ospec, 2, 0.5, anil_head=head)
        assert exp.metrics['train_loss'][it] == pytest.approx(float(out['loss'].mean()), rel=1e-4)
        assert exp.metrics['valid_loss'][it] == pytest.approx(float(out_v['loss'].mean()), rel=1e-4)
        new = mo.adam_step(body + head, [g / 2 for g in out['grad'] + out['head_grad']], state, lr=0.003)
        body, head = new[:len(body)], new[len(body):]
    got = mo.flatten([p.detach().cpu() for p in list(exp.features.parameters()) + list(exp.head.parameters())])
    assert mo.rel_l2(got, mo.flatten(body + head)) < 1e-5
    sd = torch.load(os.path.join(exp.model_path, 'head.pt'), map_location='cpu')
    assert set(sd) == {'module.weight', 'module.bias'}
    sd = torch.load(os.path.join(exp.model_path, 'features.pt'), map_location='cpu')
    assert '0.0.normalize.running_var' in sd and '0.3.conv.bias' in sd
