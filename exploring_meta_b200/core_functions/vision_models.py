"""``OmniglotCNN`` / ``MiniImagenetCNN`` / ``ConvBase`` / ``ConvBlock`` with the reference's constructor
signatures, parameter registration order and ``state_dict`` keys (``core_functions/vision_models.py:10-193``),
whose forward / backward / double-backward run on libxmeta's kernels through
``exploring_meta_b200.functional.conv_block``.

Kept from the reference: ``normalize`` (BatchNorm2d) is registered before ``conv`` (:168-185), so
``parameters()`` yields ``normalize.weight, normalize.bias, conv.weight, conv.bias`` per block; BN is used with
per-call batch statistics and updates the running buffers as a side effect (the reference never calls
``.eval()``); inits are ``uniform_`` for gamma (:175), Xavier-uniform / zero for conv and the Mini-ImageNet linear
layer (:204-207), ``normal_()`` / zero for the Omniglot linear layer (:47-49).  The sub-modules exist to hold
the parameters and buffers under the reference's names -- they are never called.
"""
import torch

from .. import functional as XF
from .._lib import XmetaError


def maml_init_(module):
    """Xavier-uniform weight, zero bias (vision_models.py:204-207)."""
    torch.nn.init.xavier_uniform_(module.weight.data, gain=1.0)
    torch.nn.init.constant_(module.bias.data, 0.0)
    return module


def _identity(x):
    return x


class ConvBlock(torch.nn.Module):
    """conv3x3(pad 1, bias) -> BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d(2, 2) when ``max_pool`` else a
    stride-2 convolution without pooling (vision_models.py:149-193)."""

    def __init__(self, in_channels, out_channels, kernel_size, max_pool=True, max_pool_factor=1.0):
        super().__init__()
        if tuple(kernel_size) != (3, 3) if not isinstance(kernel_size, int) else kernel_size != 3:
            raise XmetaError('ConvBlock: the CUDA kernels implement 3x3 convolutions only')
        stride = (int(2 * max_pool_factor), int(2 * max_pool_factor))
        if stride != (2, 2):
            raise XmetaError('ConvBlock: max_pool_factor must give a 2x2 pool / stride (got %r)' % (stride,))
        self.pooled = bool(max_pool)
        if max_pool:
            self.max_pool = torch.nn.MaxPool2d(kernel_size=stride, stride=stride, ceil_mode=False)
            stride = (1, 1)
        else:
            self.max_pool = _identity
        self.normalize = torch.nn.BatchNorm2d(out_channels, affine=True)
        torch.nn.init.uniform_(self.normalize.weight)
        self.relu = torch.nn.ReLU()
        self.conv = torch.nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=1, bias=True)
        maml_init_(self.conv)

    def forward(self, x):
        bn = self.normalize
        if not bn.training:
            raise XmetaError('ConvBlock: eval-mode BatchNorm (running statistics) is not part of the hot path; the '
                             'reference never leaves training mode')
        out, stats = XF.conv_block(x, bn.weight, bn.bias, self.conv.weight, self.conv.bias,
                                   stride=1 if self.pooled else 2, pool=self.pooled)
        if bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                XF.update_running_stats_(bn.running_mean, bn.running_var, stats, bn.momentum)
                bn.num_batches_tracked += 1
        return out


class ConvBase(torch.nn.Sequential):
    """``layers`` ConvBlocks: ``channels -> hidden`` then ``hidden -> hidden`` (vision_models.py:121-146)."""

    def __init__(self, output_size, hidden=64, channels=1, max_pool=False, layers=4, max_pool_factor=1.0):
        core = [ConvBlock(channels, hidden, (3, 3), max_pool=max_pool, max_pool_factor=max_pool_factor)]
        for _ in range(layers - 1):
            core.append(ConvBlock(hidden, hidden, kernel_size=(3, 3), max_pool=max_pool,
                                  max_pool_factor=max_pool_factor))
        super().__init__(*core)


class OmniglotCNN(torch.nn.Module):
    """vision_models.py:10-63: four stride-2 blocks on 1x28x28, mean over the 2x2 map, Linear(hidden, ways)."""

    def __init__(self, output_size=5, hidden_size=64, layers=4):
        super().__init__()
        self.hidden_size = hidden_size
        self.base = ConvBase(output_size=hidden_size, hidden=hidden_size, channels=1, max_pool=False, layers=layers)
        self.linear = torch.nn.Linear(hidden_size, output_size, bias=True)
        self.linear.weight.data.normal_()
        self.linear.bias.data.mul_(0.0)

    def forward(self, x):
        x = self.base(x.reshape(-1, 1, 28, 28))
        x = x.mean(dim=[2, 3])
        return self.linear(x)

    def get_base_representation(self, x):
        return self.base(x)

    def get_rep_layer(self, x, layer):
        if layer == -1:
            return self.linear(x.reshape(-1, 25 * self.hidden_size))
        return torch.nn.Sequential(*list(self.base.children())[:layer])(x)


class MiniImagenetCNN(torch.nn.Module):
    """vision_models.py:66-118: four pooled blocks on 3x84x84, NCHW flatten of the 5x5 map, Linear(25*hidden, ways)."""

    def __init__(self, output_size, hidden_size=32, layers=4):
        super().__init__()
        self.base = ConvBase(output_size=hidden_size, hidden=hidden_size, channels=3, max_pool=True, layers=layers,
                             max_pool_factor=4 // layers)
        self.linear = torch.nn.Linear(25 * hidden_size, output_size, bias=True)
        maml_init_(self.linear)
        self.hidden_size = hidden_size

    def forward(self, x):
        x = self.base(x)
        return self.linear(x.reshape(-1, 25 * self.hidden_size))

    def get_base_representation(self, x):
        return self.base(x)

    def get_rep_layer(self, x, layer):
        if layer == -1:
            return self.linear(x.reshape(-1, 25 * self.hidden_size))
        return torch.nn.Sequential(*list(self.base.children())[:layer])(x)


def net_spec_of(module, ways=None):
    """The ``NetSpec`` (exploring_meta_b200/spec.py) of one of the models above, or None when the module is
    not one the task-batched engine covers."""
    from ..spec import NetSpec
    if isinstance(module, MiniImagenetCNN):
        blocks = list(module.base.children())
        return NetSpec(3, 84, 84, module.hidden_size, module.linear.out_features, len(blocks), True, 'flatten')
    if isinstance(module, OmniglotCNN):
        blocks = list(module.base.children())
        if len(blocks) != 4:
            return None        # mean over the final map is part of the spec only for the 2x2 case of 4 layers
        return NetSpec(1, 28, 28, module.hidden_size, module.linear.out_features, 4, False, 'mean')
    return None
