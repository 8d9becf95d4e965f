"""Minimal ``Experiment`` base class with the reference's logging / checkpoint surface
(``utils/experiment.py:12-95``): metrics dict appended by ``log_metrics``, ``metrics.json`` / ``logger.json`` dumps,
``state_dict`` checkpoints under ``<model_path>/model_checkpoints``.  wandb and torchsummary are optional extras of
the reference and are not used here."""
import datetime
import json
import os

import numpy as np
import torch
import torch.distributed as dist


def _writer():
    """Under torchrun only rank 0 creates result directories and writes checkpoints / logs (every rank holds the same
    parameters: replicated Adam on identical sums)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank() == 0
    return int(os.environ.get('RANK', '0')) == 0


class Experiment:
    def __init__(self, algo, dataset, params, path='', use_wandb=False):
        params['algo'] = algo
        params['dataset'] = dataset
        self.params = params
        if 'seed' not in params:
            self.params.update({'seed': 42})
        seed = self.params['seed']
        self.logger = {'config': self.params, 'date': datetime.datetime.now().strftime('%d_%m_%Hh%M'),
                       'model_id': str(seed) + '_' + str(np.random.randint(1, 9999))}
        self.metrics = {}
        if path and not os.path.exists(path):
            os.makedirs(path, exist_ok=True)
        self.model_path = path + algo + '_' + dataset + '_' + self.logger['date'] + '_' + self.logger['model_id']
        if _writer():
            os.makedirs(self.model_path + '/model_checkpoints', exist_ok=True)
        self._use_wandb = False

    def log_model(self, model, device, input_shape=None, name='model'):
        if not _writer():
            return
        info = str(model)
        with open(self.model_path + '/' + name + '.summary', 'w') as file:
            file.write(info)

    def log_metrics(self, metrics, step=None):
        for key, value in metrics.items():
            self.metrics.setdefault(key, []).append(value)

    def save_logs_to_file(self):
        if not _writer():
            return
        print('Saving metrics...')
        with open(self.model_path + '/metrics.json', 'w') as fp:
            json.dump(self.metrics, fp)
        print('Saving logger...')
        with open(self.model_path + '/logger.json', 'w') as fp:
            json.dump(self.logger, fp, sort_keys=True, indent=4)

    def save_model(self, model, name='model'):
        if not _writer():
            return
        print('Saving ' + name + '...')
        torch.save(model.state_dict(), self.model_path + '/' + name + '.pt')

    def save_model_checkpoint(self, model, epoch):
        self.save_model(model, name='/model_checkpoints/model_' + epoch)

    def save_acc_matrix(self, acc_matrix):
        if not _writer():
            return
        print('Saving accuracy matrix..')
        print(acc_matrix)
        np.savetxt(self.model_path + '/acc_matrix.out', acc_matrix, fmt='%1.2f')
