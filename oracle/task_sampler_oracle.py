"""CPU restatement of the on-device task sampler (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's CPU leg
may import anything under oracle/).

What it restates: the task construction of the reference's data pipeline, ``utils/data_pre.py:28-37`` (Omniglot)
and ``:79-85`` (Mini-ImageNet) -- learn2learn's transform chain
``NWays -> KShots(2k) -> LoadData -> RemapLabels -> ConsecutiveLabels (-> RandomClassRotation)``:
``ways`` distinct classes, ``2k`` distinct items of each, samples grouped by class, labels remapped to
``0..ways-1``, for Omniglot one rotation out of {0, 90, 180, 270} degrees per class of the task; pixel transforms
``1 - x/255`` (Omniglot, ``data_pre.py:18-22``) or raw 0..255 floats (Mini-ImageNet).  learn2learn draws with
Python's ``random``; that stream cannot be reproduced on a GPU, so parity with the reference is semantic
(the properties ``tests/test_task_sampler.py`` checks) while parity between this oracle and the CUDA kernel
(``exploring_meta_b200/csrc/sampler.cu``) is bit-exact: both use splitmix64 of (seed, task number, draw counter),
multiply-shift range reduction and sequential rejection for distinctness.

parity unpinned: the reference ships no fixtures for this path.
"""
import numpy as np

_M = (1 << 64) - 1


def splitmix64(x):
    x &= _M
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & _M
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & _M
    x ^= x >> 31
    return x


def draw(seed, task, ctr, n):
    """Draw number ``ctr`` of task ``task``: uniform integer in [0, n)."""
    h = splitmix64((seed ^ ((task * 0x9E3779B97F4A7C15) & _M) ^ ((ctr * 0xD1B54A32D192ED03) & _M)) & _M)
    return ((h >> 32) * n) >> 32


def sample_task_indices(seed, task, class_start, ways, shots2, rotate):
    """Returns (classes [ways], items [ways*shots2], quarter_turns [ways]) of global task number ``task``."""
    num_classes = len(class_start) - 1
    ctr = 0
    classes = []
    for _ in range(ways):                       # NWays
        while True:
            c = draw(seed, task, ctr, num_classes)
            ctr += 1
            if c not in classes:
                break
        classes.append(c)
    items = []
    for c in classes:                           # KShots(2k), without replacement
        lo, cnt = int(class_start[c]), int(class_start[c + 1] - class_start[c])
        mine = []
        for _ in range(shots2):
            while True:
                it = lo + draw(seed, task, ctr, cnt)
                ctr += 1
                if it not in mine:
                    break
            mine.append(it)
        items += mine
    rots = []
    for _ in range(ways):                       # RandomClassRotation
        if rotate:
            rots.append(draw(seed, task, ctr, 4))
            ctr += 1
        else:
            rots.append(0)
    return classes, items, rots


def sample_tasks(data, class_start, tasks, ways, shots2, seed, first_task, rotate=False, scale=1.0, offset=0.0):
    """data: uint8 [items][C][H][W].  Returns x float32 [tasks][ways*shots2][C][H][W], y int64 [tasks][ways*shots2],
    items int32 [tasks][ways*shots2], classes int32 [tasks][ways]."""
    data = np.asarray(data)
    per = ways * shots2
    x = np.empty((tasks, per) + data.shape[1:], dtype=np.float32)
    y = np.tile(np.repeat(np.arange(ways, dtype=np.int64), shots2), (tasks, 1))     # RemapLabels + ConsecutiveLabels
    items = np.empty((tasks, per), dtype=np.int32)
    classes = np.empty((tasks, ways), dtype=np.int32)
    for t in range(tasks):
        cl, it, rots = sample_task_indices(seed, first_task + t, class_start, ways, shots2, rotate)
        classes[t], items[t] = cl, it
        for s, item in enumerate(it):
            img = np.rot90(data[item], k=rots[s // shots2], axes=(1, 2))             # counter-clockwise quarter turns
            # same arithmetic as the kernel: one fused multiply-add in fp32
            x[t, s] = (np.float64(np.float32(scale)) * img.astype(np.float64) + np.float64(np.float32(offset))).astype(np.float32)
    return x, y, items, classes
