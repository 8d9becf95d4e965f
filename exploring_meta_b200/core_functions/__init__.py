"""Host-side mirror of the reference's ``core_functions`` package for the vision hot path (and, in ``rl`` /
``policies``, the MAML-TRPO policy path):
``maml.MAML`` (clone / adapt), ``vision.fast_adapt / accuracy / evaluate`` and the ``vision_models`` CNNs, all
running on libxmeta's sm_100a kernels."""
from .maml import MAML                                              # noqa: F401
from .vision import accuracy, evaluate, fast_adapt                  # noqa: F401
from .vision_models import ConvBase, ConvBlock, MiniImagenetCNN, OmniglotCNN   # noqa: F401
