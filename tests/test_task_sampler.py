"""On-device task sampler: the oracle's restatement of the reference's task construction
(utils/data_pre.py:28-37, 79-85) has the properties of a learn2learn task; the host class on the C-ABI emulator and
the CUDA kernel reproduce the oracle bit for bit (integer / byte work: exact)."""
import numpy as np
import pytest
import torch

from oracle import task_sampler_oracle as tso


def _dataset(num_classes, per_class, shape, seed=0, ragged=False):
    rng = np.random.RandomState(seed)
    counts = [per_class + (rng.randint(0, 4) if ragged else 0) for _ in range(num_classes)]
    labels = np.repeat(np.arange(num_classes) * 7 + 3, counts)                 # arbitrary class ids
    data = rng.randint(0, 256, size=(labels.size,) + shape, dtype=np.uint8)
    perm = rng.permutation(labels.size)                                        # items arrive unsorted
    return data[perm], labels[perm]


def test_oracle_tasks_look_like_learn2learn_tasks():
    ways, shots = 5, 2
    data, labels = _dataset(12, 6, (1, 8, 8))
    order = np.argsort(labels, kind='stable')
    data, labels = data[order], labels[order]
    cs = np.concatenate([[0], np.cumsum(np.bincount((labels - 3) // 7))]).astype(np.int32)
    x, y, items, classes = tso.sample_tasks(data, cs, 40, ways, 2 * shots, seed=123, first_task=5, rotate=True,
                                            scale=-1 / 255.0, offset=1.0)
    assert x.shape == (40, 20, 1, 8, 8) and y.shape == (40, 20)
    for t in range(40):
        assert len(set(classes[t].tolist())) == ways                           # NWays: distinct classes
        assert y[t].tolist() == sorted(y[t].tolist()) and y[t, ::2 * shots].tolist() == list(range(ways))
        for w in range(ways):
            it = items[t, w * 2 * shots:(w + 1) * 2 * shots]
            assert len(set(it.tolist())) == 2 * shots                          # KShots: without replacement
            assert all(cs[classes[t, w]] <= i < cs[classes[t, w] + 1] for i in it)   # ... from the drawn class
        # every image is a quarter-turn rotation of its item under the pixel transform, one rotation per class
        for w in range(ways):
            ks = set()
            for s in range(w * 2 * shots, (w + 1) * 2 * shots):
                src = 1.0 - data[items[t, s]].astype(np.float32) / 255.0
                k = [k for k in range(4) if np.allclose(np.rot90(src, k, axes=(1, 2)), x[t, s], atol=1e-6)]
                assert k
                ks.add(tuple(k))
            assert len(ks) == 1 or any(len(k) > 1 for k in ks)
    # a batch is a pure function of (seed, first task number)
    x2, y2, items2, _ = tso.sample_tasks(data, cs, 3, ways, 2 * shots, seed=123, first_task=7, rotate=True,
                                         scale=-1 / 255.0, offset=1.0)
    assert np.array_equal(items2, items[2:5]) and np.array_equal(x2, x[2:5])
    # all classes / rotations get used
    assert len(set(classes.flatten().tolist())) == 12


@pytest.mark.parametrize('dev', ['emulator', pytest.param('cuda', marks=pytest.mark.gpu)])
@pytest.mark.parametrize('case', [
    dict(num_classes=30, per_class=20, shape=(1, 28, 28), ways=20, shots=5, rotate=True, transform='omni', tasks=6),
    dict(num_classes=9, per_class=14, shape=(3, 84, 84), ways=5, shots=5, rotate=False, transform='raw', tasks=4),
    dict(num_classes=7, per_class=3, shape=(2, 5, 5), ways=5, shots=1, rotate=True, transform='raw', tasks=33, ragged=True),
])
def test_sampler_matches_oracle_bit_for_bit(dev, case, monkeypatch):
    from exploring_meta_b200.utils import device_tasks as dt
    if dev == 'emulator':
        import cabi_emulator
        from exploring_meta_b200 import _lib
        monkeypatch.setattr(_lib, '_lib', cabi_emulator.EmulatedLib())
        device = 'cpu'
    else:
        device = 'cuda'
    data, labels = _dataset(case['num_classes'], case['per_class'], case['shape'], seed=1, ragged=case.get('ragged', False))
    tr = dt.OMNIGLOT_TRANSFORM if case['transform'] == 'omni' else dt.RAW_TRANSFORM
    s = dt.DeviceTaskSampler(torch.from_numpy(data), labels, case['ways'], case['shots'], rotate=case['rotate'],
                             transform=tr, seed=99, device=device)
    x, y, items, classes = s.sample_batch(case['tasks'], return_indices=True)
    x_next, _ = s.sample_batch(2)                                              # the counter continues
    ref = tso.sample_tasks(s.data.cpu().numpy(), s.class_start.cpu().numpy(), case['tasks'] + 2, case['ways'],
                           2 * case['shots'], 99, 0, rotate=case['rotate'], scale=tr[0], offset=tr[1])
    n = case['tasks']
    assert np.array_equal(items.cpu().numpy(), ref[2][:n])
    assert np.array_equal(classes.cpu().numpy(), ref[3][:n])
    assert np.array_equal(y.cpu().numpy(), ref[1][:n])
    assert np.array_equal(x.cpu().numpy(), ref[0][:n])
    assert np.array_equal(x_next.cpu().numpy(), ref[0][n:])
    xs, ys = s.sample()
    assert xs.shape == x.shape[1:] and ys.tolist() == y[0].tolist()


def test_sampler_feeds_the_engine(emulated_lib):
    """A sampled batch has exactly the layout MamlEngine consumes (support = even rows, query = odd rows)."""
    from exploring_meta_b200 import engine as eng, spec as pspec
    from exploring_meta_b200.utils import device_tasks as dt
    data, labels = _dataset(8, 6, (1, 14, 14), seed=2)
    s = dt.DeviceTaskSampler(torch.from_numpy(data), labels, 4, 1, rotate=True, transform=dt.OMNIGLOT_TRANSFORM,
                             seed=5, device='cpu')
    x, y = s.sample_batch(2)
    spec = pspec.NetSpec(1, 14, 14, 8, 4, 3, False, 'mean')
    e = eng.MamlEngine(spec, 2, 1, 1, 0.4, mode='second', device='cpu')
    e.run(x, y, pspec.init_flat_params(spec))
    assert torch.isfinite(e.grad).all() and torch.isfinite(e.loss).all()
    assert y[:, 0::2].tolist() == y[:, 1::2].tolist()            # every class has k support and k query rows
