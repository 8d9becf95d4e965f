"""On-device task sampler: the step BEFORE the hot path.

``DeviceTaskSampler`` stands in for the learn2learn ``TaskDataset`` objects the reference builds in
``utils/data_pre.py:16-112`` (``get_omniglot`` / ``get_mini_imagenet``): the image set lives in HBM as uint8, and
``sample_batch(tasks)`` draws ``tasks`` few-shot tasks in one kernel (``xm_sample_tasks``) -- N-way / K-shot index
sampling without replacement, label remapping, per-class Omniglot rotations, the pixel transform -- straight into the
``[tasks, 2*k*w, C, H, W]`` / ``[tasks, 2*k*w]`` layout ``MamlEngine`` consumes, so no image ever crosses PCIe during
training.  ``sample()`` returns one task as ``(data, labels)`` like ``TaskDataset.sample()`` (``vision/maml_vision.py:105``,
``core_functions/vision.py:32``), which makes the object a drop-in ``train_tasks`` for ``fast_adapt`` / ``evaluate``.
Draws are counter based: task number n of a sampler is a pure function of (seed, n), so ranks shard a global
meta-batch by task number without communication.
"""
import ctypes

import torch

from .. import _lib
from .._lib import XmSampleArgs

OMNIGLOT_TRANSFORM = (-1.0 / 255.0, 1.0)      # ToTensor then 1 - x  (utils/data_pre.py:18-22)
RAW_TRANSFORM = (1.0, 0.0)                    # learn2learn Mini-ImageNet: raw 0..255 floats


class DeviceTaskSampler:
    def __init__(self, data_u8, labels, ways, shots, rotate=False, transform=RAW_TRANSFORM, seed=0, device='cuda'):
        """data_u8: uint8 tensor [items, C, H, W]; labels: integer class of every item (any integers).
        Items are re-ordered by class once; classes with fewer than 2*shots items are dropped (KShots needs them)."""
        self.device = torch.device(device)
        self.ways, self.shots, self.rotate = int(ways), int(shots), bool(rotate)
        self.scale, self.offset = float(transform[0]), float(transform[1])
        self.seed = int(seed) & ((1 << 64) - 1)
        labels = torch.as_tensor(labels).to(torch.int64).cpu()
        order = torch.argsort(labels, stable=True)
        uniq, counts = torch.unique_consecutive(labels[order], return_counts=True)
        keep = counts >= 2 * self.shots
        starts = torch.cumsum(counts, 0) - counts
        idx = torch.cat([order[s:s + c] for s, c, k in zip(starts.tolist(), counts.tolist(), keep.tolist()) if k])
        kept = counts[keep]
        if kept.numel() < self.ways:
            raise ValueError('only %d classes with >= %d items for %d-way tasks' % (kept.numel(), 2 * self.shots, self.ways))
        self.class_ids = uniq[keep]
        cs = torch.zeros(kept.numel() + 1, dtype=torch.int32)
        cs[1:] = torch.cumsum(kept, 0).to(torch.int32)
        data_u8 = torch.as_tensor(data_u8)
        assert data_u8.dtype == torch.uint8 and data_u8.dim() == 4
        self.data = data_u8[idx].contiguous().to(self.device)
        self.class_start = cs.to(self.device)
        self.shape = tuple(self.data.shape[1:])
        self.next_task = 0
        self.lib = _lib.load()

    def sample_batch(self, tasks, first_task=None, return_indices=False):
        """``tasks`` tasks numbered first_task .. first_task + tasks - 1 (default: continue this sampler's counter)."""
        if first_task is None:
            first_task = self.next_task
            self.next_task += tasks
        per = 2 * self.shots * self.ways
        x = torch.empty((tasks, per) + self.shape, dtype=torch.float32, device=self.device)
        y = torch.empty((tasks, per), dtype=torch.int64, device=self.device)
        items = torch.empty((tasks, per), dtype=torch.int32, device=self.device) if return_indices else None
        classes = torch.empty((tasks, self.ways), dtype=torch.int32, device=self.device) if return_indices else None
        a = XmSampleArgs()
        a.tasks, a.ways, a.shots2 = tasks, self.ways, 2 * self.shots
        a.channels, a.height, a.width = self.shape
        a.num_classes, a.rotate = self.class_start.numel() - 1, int(self.rotate)
        a.seed, a.first_task = self.seed, int(first_task)
        a.data, a.class_start = self.data.data_ptr(), self.class_start.data_ptr()
        a.scale, a.offset = self.scale, self.offset
        a.x, a.y = x.data_ptr(), y.data_ptr()
        a.items = items.data_ptr() if return_indices else None
        a.classes = classes.data_ptr() if return_indices else None
        stream = torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0
        _lib.check(self.lib.xm_sample_tasks(ctypes.byref(a), stream), 'xm_sample_tasks')
        return (x, y, items, classes) if return_indices else (x, y)

    def sample(self):
        """One task, like ``TaskDataset.sample()``: ``(data [2kw, C, H, W], labels [2kw])``."""
        x, y = self.sample_batch(1)
        return x[0], y[0]
