"""``prepare_batch`` (reference ``utils/data_pre.py:115-129``): move a sampled task to the device, optionally
run the ANIL body over ALL rows (one BatchNorm batch), and split rows 0, 2, 4, ... (the first ``shots * ways``
even rows) into the adaptation set and every other row into the evaluation set.

The dataset builders of the reference (``get_omniglot`` / ``get_mini_imagenet``: learn2learn TaskDatasets that
download the data) are outside the hot path; ``exploring_meta_b200.synthetic.SyntheticTasks`` provides the
``sample()`` interface on seeded synthetic tensors of the same shapes.
"""
import torch


def prepare_batch(batch, shots, ways, device, features=None):
    data, labels = batch
    data, labels = data.to(device), labels.to(device)
    if features is not None:
        data = features(data)
    rows = data.size(0)
    idx = torch.arange(rows, device=data.device)
    adaptation = (idx % 2 == 0) & (idx < 2 * shots * ways)
    evaluation = ~adaptation
    return data[adaptation], labels[adaptation], data[evaluation], labels[evaluation]
