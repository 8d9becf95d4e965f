"""Loading of tests/golden/*.npz (written by tests/golden/make_golden.py from the reference's own files)
and regeneration of the matching seeded inputs."""
import os

import numpy as np
import torch

from exploring_meta_b200.synthetic import make_tasks
from oracle import maml_oracle as mo

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
IN_SHAPE = {'omni': (1, 28, 28), 'min': (3, 84, 84)}
NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith('.npz'))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.name = name
        self.algo, self.kind = (str(v) for v in z['case'])
        for k in ('ways', 'shots', 'steps', 'tasks', 'data_seed', 'model_seed'):
            setattr(self, k, int(z[k]))
        self.inner_lr = float(z['inner_lr'])
        self.z = z

    def ospec(self):
        if self.algo == 'maml':
            return mo.omniglot_spec(self.ways) if self.kind == 'omni' else mo.miniimagenet_spec(self.ways)
        if self.kind == 'omni':
            return mo.NetSpec(1, 28, 28, 32, self.ways, 4, False, 'flatten')
        return mo.NetSpec(3, 84, 84, 64, self.ways, 4, True, 'flatten')

    def inputs(self):
        """(X, Y, body/all params fp32 list, head params fp32 list or None), checked against the checksums."""
        X, Y = make_tasks(self.tasks, self.ways, self.shots, IN_SHAPE[self.kind], seed=self.data_seed)
        assert abs(X.double().sum().item() - float(self.z['x_checksum'])) <= 1e-9 * float(self.z['x_abs_checksum'])
        if self.algo == 'maml':
            params, head = mo.init_params(self.ospec(), seed=self.model_seed), None
        else:
            params, head = mo.init_anil_params(self.ospec(), seed=self.model_seed)
        chk = mo.flatten(params).double().sum().item()
        assert abs(chk - float(self.z['theta_checksum'])) <= 1e-9 * max(1.0, abs(chk))
        return X, Y, params, head

    def grad_mask(self):
        return ~mo.conv_bias_mask(self.ospec(), with_head=(self.algo == 'maml'))

    def t(self, key):
        return torch.from_numpy(np.asarray(self.z[key]))
