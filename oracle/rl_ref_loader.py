"""Import the reference's own RL hot-path files (core_functions/rl.py, core_functions/policies.py), unmodified, from
/root/reference on top of the cherry / learn2learn restatements (TEST INFRASTRUCTURE, build container only)."""
import importlib
import os
import sys
import types

from . import ref_loader

_loaded = None


def available():
    return os.path.isfile(os.path.join(ref_loader.REFERENCE_ROOT, 'core_functions', 'rl.py'))


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    ref_loader.load()                                   # learn2learn shim + bare utils / core_functions packages
    from . import cherry_shim
    cherry_shim.install()
    sys.modules['utils'].make_env = None                # rl.py:15 `from utils import make_env` (environment factory: unused here)
    runner = types.ModuleType('core_functions.runner')  # rl.py:16 (rollout collection: out of scope)
    runner.Runner = None
    sys.modules['core_functions.runner'] = runner
    policies = importlib.import_module('core_functions.policies')
    rl = importlib.import_module('core_functions.rl')
    _loaded = types.SimpleNamespace(rl=rl, policies=policies, cherry=cherry_shim)
    return _loaded
