#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE'S OWN FILES (build container only).

    python tests/golden/make_golden.py [case ...] # needs /root/reference; no names = every case

For every case below the train half of one outer iteration (vision/maml_vision.py:95-112 or
vision/anil_vision.py:109-122) is replayed through oracle/reference_run.py -- the reference's unmodified
``fast_adapt`` / ``prepare_batch`` / model constructors / ``MAML`` subclass on top of the learn2learn
restatement -- in fp64 and fp32, on the seeded synthetic tasks of exploring_meta_b200/synthetic.py.  Before
anything is written the self-contained restatement oracle/maml_oracle.py must reproduce the fp64 run
to <= 1e-11 (rel-L2 of the summed meta-gradient, per-task query loss; exact correct counts).

The fixture stores the fp64 outputs (query loss, correct count; adapted weights of task 0 and the summed
meta-gradient rounded to fp32 to keep the files small -- the tolerances are >= 1e-4), the fp32-vs-fp64
deviation of the reference itself (``e_ref``, the yardstick of the tolerance contract) and checksums of the inputs, so a test on any machine can tell whether it regenerated the same
inputs.  Regenerate only when the case list changes; the files are committed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from exploring_meta_b200.synthetic import make_tasks      # noqa: E402
from oracle import maml_oracle as mo                       # noqa: E402
from oracle import reference_run as rr                     # noqa: E402

# name -> (algo, dataset kind, ways, shots, steps, inner_lr, tasks, data seed)
CASES = {
    # BASELINE.json configs[0]: MAML Omniglot 5-way 1-shot, 64-filter CNN, 1 inner step, second-order
    'maml_omni_5w1s_t1': ('maml', 'omni', 5, 1, 1, 0.5, 4, 0),
    # configs[3] shape (20-way 5-shot Omniglot), 1 task
    'maml_omni_20w5s_t1': ('maml', 'omni', 20, 5, 1, 0.5, 1, 1),
    # configs[1] network (MiniImagenetCNN, 32 filters) at the calm inner lr, 2 steps
    'maml_min_5w1s_t2_calm': ('maml', 'min', 5, 1, 2, 0.001, 2, 3),
    # configs[1] exactly as named but for the task count: 5-way 5-shot, 5 steps, lr 0.5 (chaotic regime:
    # the fixture records how far the reference's own fp32 run is from its fp64 run)
    'maml_min_5w5s_t5_headline': ('maml', 'min', 5, 5, 5, 0.5, 1, 0),
    # configs[2]: ANIL Mini-ImageNet 5-way 5-shot, 64-filter body, head-only adaptation, 1 step
    'anil_min_5w5s_t1': ('anil', 'min', 5, 5, 1, 0.5, 2, 6),
    'anil_omni_5w1s_t2': ('anil', 'omni', 5, 1, 2, 0.5, 3, 7),
    # configs[1] EXACTLY (5-way 5-shot => S = 25 support rows, T = 5 inner steps) at the calm inner lr, 2 tasks:
    # the tight-tolerance twin of the chaotic headline case.  Even at this lr the S = 25, T = 5 shape is rarely calm:
    # over data seeds 100..115 the reference's own fp32 run leaves the fp64 run's ReLU / max-pool decisions in 11 of
    # 16 seeds (e_ref 1e-4 .. 4e-2) and the CUDA path in 10 of 16, often on the same decision
    # (scripts/diag_flip_rate.py, profiles/r02_flip_rate.txt).  Seed 108 is one where both stay on them.
    'maml_min_5w5s_t5_calm': ('maml', 'min', 5, 5, 5, 0.001, 1, 108),
    # configs[0] at its full meta-batch: 32 tasks
    'maml_omni_5w1s_t1_b32': ('maml', 'omni', 5, 1, 1, 0.5, 32, 8),
}

IN_SHAPE = {'omni': (1, 28, 28), 'min': (3, 84, 84)}


def ospec_of(algo, kind, ways):
    if algo == 'maml':
        return mo.omniglot_spec(ways) if kind == 'omni' else mo.miniimagenet_spec(ways)
    if kind == 'omni':
        return mo.NetSpec(1, 28, 28, 32, ways, 4, False, 'flatten')
    return mo.NetSpec(3, 84, 84, 64, ways, 4, True, 'flatten')


def run_reference(algo, kind, ways, shots, steps, lr, X, Y, dtype):
    if algo == 'maml':
        model = rr.build_model(kind, ways, seed=42, dtype=dtype)
        out = rr.maml_iteration(model, X.to(dtype), Y, ways, shots, steps, lr)
        out['params'] = [p.detach().clone() for p in model.parameters()]
        return out
    features, head = rr.build_anil(kind, ways, seed=42, dtype=dtype)
    out = rr.anil_iteration(features, head, X.to(dtype), Y, ways, shots, steps, lr)
    out['params'] = [p.detach().clone() for p in features.parameters()]
    out['head_params'] = [p.detach().clone() for p in head.parameters()]
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    wanted = sys.argv[1:] or list(CASES)
    for name in wanted:
        algo, kind, ways, shots, steps, lr, tasks, seed = CASES[name]
        X, Y = make_tasks(tasks, ways, shots, IN_SHAPE[kind], seed=seed)
        r64 = run_reference(algo, kind, ways, shots, steps, lr, X, Y, torch.float64)
        r32 = run_reference(algo, kind, ways, shots, steps, lr, X, Y, torch.float32)
        ospec = ospec_of(algo, kind, ways)
        S = ways * shots
        # ---- the restatement must reproduce the reference-file run --------------------------------------
        if algo == 'maml':
            p32 = mo.init_params(ospec, seed=42)
            for a, b in zip(p32, r32['params']):
                assert torch.equal(a, b), 'init_params does not reproduce the reference constructor'
            o64 = mo.meta_iteration([p.double() for p in p32], X.double(), Y, ospec, steps, lr)
        else:
            b32, h32 = mo.init_anil_params(ospec, seed=42)
            for a, b in zip(b32 + h32, r32['params'] + r32['head_params']):
                assert torch.equal(a, b), 'init_anil_params does not reproduce the reference constructors'
            o64 = mo.meta_iteration([p.double() for p in b32], X.double(), Y, ospec, steps, lr,
                                    anil_head=[p.double() for p in h32])
        g_ref, g_or = mo.flatten(r64['grad']), mo.flatten(o64['grad'])
        mask = ~mo.conv_bias_mask(ospec, with_head=(algo == 'maml'))
        d = mo.rel_l2(g_or[mask], g_ref[mask])
        assert d <= 1e-11, '%s: restatement vs reference files, meta-grad rel-L2 %.3e' % (name, d)
        assert torch.allclose(o64['loss'], r64['loss'], rtol=1e-12, atol=1e-13)
        assert [int(round(a * S)) for a in r64['acc'].tolist()] == o64['correct'].tolist()
        if algo == 'anil':
            dh = mo.rel_l2(mo.flatten(o64['head_grad']), mo.flatten(r64['head_grad']))
            assert dh <= 1e-11, '%s: head grad %.3e' % (name, dh)
        # ---- the reference's own fp32 deviation ----------------------------------------------------------
        g32 = mo.flatten(r32['grad'])
        e_ref = mo.rel_l2(g32[mask], g_ref[mask])
        fix = {
            'case': np.array([algo, kind]), 'ways': ways, 'shots': shots, 'steps': steps, 'inner_lr': lr,
            'tasks': tasks, 'data_seed': seed, 'model_seed': 42,
            'x_checksum': X.double().sum().item(), 'x_abs_checksum': X.double().abs().sum().item(),
            'theta_checksum': mo.flatten(r64['params']).sum().item(),
            'loss64': r64['loss'].numpy(), 'loss32': r32['loss'].numpy(),
            'correct': np.array([int(round(a * S)) for a in r64['acc'].tolist()], dtype=np.int64),
            'correct32': np.array([int(round(a * S)) for a in r32['acc'].tolist()], dtype=np.int64),
            'grad64_as_f32': g_ref.float().numpy(), 'e_ref_grad': e_ref,
            'conv_bias_grad_absmax32': float(g32[~mask].abs().max()) if (~mask).any() else 0.0,
        }
        if algo == 'maml':
            th64, th32 = mo.flatten(r64['adapted'][0]), mo.flatten(r32['adapted'][0])
            fix['adapted0_64_as_f32'] = th64.float().numpy()
            fix['e_ref_adapted0'] = mo.rel_l2(th32[mask], th64[mask])
        else:
            fix['head_grad64_as_f32'] = mo.flatten(r64['head_grad']).float().numpy()
            fix['e_ref_head_grad'] = mo.rel_l2(mo.flatten(r32['head_grad']), mo.flatten(r64['head_grad']))
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **fix)
        print('%-28s restatement-vs-reference %.1e | e_ref(grad) %.2e | loss %s | correct %s | %.0f KB'
              % (name, d, e_ref, np.round(fix['loss64'], 5).tolist(), fix['correct'].tolist(),
                 os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
