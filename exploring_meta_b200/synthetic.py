"""Seeded synthetic few-shot tasks of the reference's shapes.

A task is what ``train_tasks.sample()`` returns in the reference (``vision/maml_vision.py:105``):
``(data [2*k*w, C, H, W] float32, labels [2*k*w] int64)`` with samples grouped by class, 2k
consecutive rows per class, so that ``prepare_batch`` (``utils/data_pre.py:115-129``) puts k rows of
every class into the support set (even rows) and k into the query set (odd rows).

Images are ``eps + 0.5 * prototype[label]`` with ``eps, prototype ~ N(0, 1)``: one prototype per
class per task gives the inner loop a learnable signal.  Generation is on the CPU generator so the
same seed yields the same bits in the build container, on the GPU box and in the fixtures.
"""
import torch


def task_labels(ways, shots):
    return torch.arange(ways, dtype=torch.int64).repeat_interleave(2 * shots)


def make_tasks(tasks, ways, shots, in_shape, seed, dtype=torch.float32):
    """Returns ``x [tasks, 2*k*w, C, H, W]`` and ``y [tasks, 2*k*w]`` on the CPU."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    y = task_labels(ways, shots)
    proto = torch.randn((tasks, ways) + tuple(in_shape), generator=gen, dtype=torch.float32)
    x = torch.randn((tasks, y.numel()) + tuple(in_shape), generator=gen, dtype=torch.float32)
    x.add_(proto[:, y], alpha=0.5)
    return x.to(dtype), y.unsqueeze(0).repeat(tasks, 1).contiguous()


class SyntheticTasks:
    """Stand-in for a learn2learn ``TaskDataset``: ``sample()`` returns one task, as the reference's
    drivers and ``evaluate`` expect (``core_functions/vision.py:32``)."""

    def __init__(self, ways, shots, in_shape, seed=0):
        self.ways, self.shots, self.in_shape = ways, shots, tuple(in_shape)
        self._seed = int(seed)
        self._count = 0

    def sample(self):
        x, y = make_tasks(1, self.ways, self.shots, self.in_shape, self._seed + self._count)
        self._count += 1
        return x[0], y[0]

    def sample_batch(self, tasks):
        x, y = make_tasks(tasks, self.ways, self.shots, self.in_shape, self._seed + self._count)
        self._count += tasks
        return x, y


def get_tasks(dataset, ways, shots, seed=0):
    """Stand-in for the reference's ``get_omniglot`` / ``get_mini_imagenet`` (``utils/data_pre.py:16-112``, learn2learn
    datasets that need a download): three independent synthetic task streams ``(train, valid, test)``."""
    shape = (1, 28, 28) if dataset == 'omni' else (3, 84, 84)
    return tuple(SyntheticTasks(ways, shots, shape, seed=seed + 1_000_003 * k) for k in range(3))


def make_replays(tasks, episodes=20, horizon=100, seed=0, dtype=torch.float32):
    """Synthetic Particles2D-style rollouts for the MAML-TRPO policy path: per task a
    support and a query replay of ``episodes * horizon`` transitions -- states ~ 0.3 N(0, 1) random walks, actions
    ~ 0.1 N(0, 1), reward = -||s' - goal|| with goal ~ U[-0.5, 0.5]^2, ``done`` at the end of every episode.
    Returns a list (tasks) of [support, query] dicts with keys states / actions / rewards / dones / next_states."""
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    n = episodes * horizon
    out = []
    for _ in range(tasks):
        goal = torch.rand(2, generator=gen, dtype=torch.float32) - 0.5
        pair = []
        for _k in range(2):
            s = 0.3 * torch.randn(n, 2, generator=gen, dtype=torch.float32)   # fp32 draws whatever the default dtype
            a = 0.1 * torch.randn(n, 2, generator=gen, dtype=torch.float32)
            ns = s + a
            r = -(ns - goal).norm(dim=1, keepdim=True)
            d = torch.zeros(n, 1, dtype=torch.float32)
            d[horizon - 1::horizon] = 1.0
            pair.append({'states': s.to(dtype), 'actions': a.to(dtype), 'rewards': r.to(dtype), 'dones': d.to(dtype),
                         'next_states': ns.to(dtype)})
        out.append(pair)
    return out
