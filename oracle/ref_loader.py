"""Import the reference's own hot-path files, unmodified, from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: ``/root/reference`` does not exist on the GPU box, so nothing
in the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that are skipped when the
reference tree is absent.

Recipe: (1) install the learn2learn restatement under the names the reference imports;
(2) register bare ``utils`` / ``core_functions`` packages whose ``__path__`` points into the
reference so that sub-modules load WITHOUT running the reference ``__init__.py`` files (those pull
in cherry / gym / metaworld / matplotlib, none of which is installed).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('XM_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'core_functions', 'vision.py'))


_loaded = None


def load():
    """Returns a namespace with the reference's fast_adapt / accuracy / evaluate / prepare_batch /
    OmniglotCNN / MiniImagenetCNN / ConvBase / ConvBlock and the shimmed MAML wrapper subclass."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    from . import l2l_shim
    l2l_shim.install()
    for pkg in ('utils', 'core_functions'):
        if pkg in sys.modules and getattr(sys.modules[pkg], '__file__', None):
            raise RuntimeError('a real package named %r is already imported' % pkg)
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
        sys.modules[pkg] = m
    vision = importlib.import_module('core_functions.vision')
    models = importlib.import_module('core_functions.vision_models')
    data_pre = importlib.import_module('utils.data_pre')
    maml_mod = importlib.import_module('core_functions.maml')
    ns = types.SimpleNamespace(
        fast_adapt=vision.fast_adapt, accuracy=vision.accuracy, evaluate=vision.evaluate,
        prepare_batch=data_pre.prepare_batch,
        OmniglotCNN=models.OmniglotCNN, MiniImagenetCNN=models.MiniImagenetCNN,
        ConvBase=models.ConvBase, ConvBlock=models.ConvBlock,
        MAML=maml_mod.MAML, l2l=l2l_shim)
    _loaded = ns
    return ns
