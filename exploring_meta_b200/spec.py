"""Network shape description shared by the host-side modules.

Mirrors the constructors of ``core_functions/vision_models.py``: ``ConvBase`` (:121-146) is
``layers`` ConvBlocks (first ``channels -> hidden``, rest ``hidden -> hidden``); with ``max_pool=True``
each block is conv(stride 1) + MaxPool2d(2, 2), otherwise conv(stride 2) without pooling (:158-167).
``MiniImagenetCNN`` (:91-110) flattens the 5x5 map in NCHW order into ``Linear(25*hidden, ways)``;
``OmniglotCNN`` (:38-55) averages the 2x2 map into ``Linear(hidden, ways)``.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class NetSpec:
    in_c: int
    in_h: int
    in_w: int
    hidden: int
    ways: int
    layers: int = 4
    pool: bool = True
    head: str = 'flatten'      # 'flatten' | 'mean' | 'none' (ANIL body: features only)

    # ---- geometry -------------------------------------------------------------------------
    def block_dims(self):
        """Per block: (cin, hin, win, hz, wz, hp, wp)."""
        out, cin, h, w = [], self.in_c, self.in_h, self.in_w
        for _ in range(self.layers):
            if self.pool:
                hz, wz = h, w
                hp, wp = hz // 2, wz // 2
            else:
                hz, wz = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
                hp, wp = hz, wz
            out.append((cin, h, w, hz, wz, hp, wp))
            cin, h, w = self.hidden, hp, wp
        return out

    def out_hw(self):
        d = self.block_dims()[-1]
        return d[5], d[6]

    def feat_dim(self):
        h, w = self.out_hw()
        return self.hidden if self.head == 'mean' else self.hidden * h * w

    # ---- flat parameter vector in module.parameters() order -----------------------------------
    def param_shapes(self):
        shapes, cin = [], self.in_c
        for _ in range(self.layers):
            shapes += [(self.hidden,), (self.hidden,), (self.hidden, cin, 3, 3), (self.hidden,)]
            cin = self.hidden
        if self.head != 'none':
            shapes += [(self.ways, self.feat_dim()), (self.ways,)]
        return shapes

    def param_offsets(self):
        offs, o = [], 0
        for shp in self.param_shapes():
            offs.append(o)
            n = 1
            for s in shp:
                n *= s
            o += n
        return offs, o

    @property
    def num_params(self):
        return self.param_offsets()[1]

    def flops_per_train_task(self, shots, steps):
        """Algorithmic conv+linear FLOPs (2/MAC) of one second-order MAML train task:
        per support image per step 4*M1 + 9*sum(M_l>=2), per query image 2*M1 + 3*sum(M_l>=2)."""
        m = []
        for (cin, _h, _w, hz, wz, _hp, _wp) in self.block_dims():
            m.append(hz * wz * self.hidden * cin * 9)
        rest = sum(m[1:]) + (self.feat_dim() * self.ways if self.head != 'none' else 0)
        s = self.ways * shots
        macs = s * (steps * (4 * m[0] + 9 * rest) + (2 * m[0] + 3 * rest))
        return 2 * macs


def miniimagenet_spec(ways=5, hidden=32, layers=4):
    return NetSpec(3, 84, 84, hidden, ways, layers, True, 'flatten')


def omniglot_spec(ways=5, hidden=64, layers=4):
    return NetSpec(1, 28, 28, hidden, ways, layers, False, 'mean')


def anil_body_spec(dataset, ways=5):
    """``vision/anil_vision.py:86-89``: Omniglot body hidden=32 stride-2, Mini-ImageNet body ConvBase
    default hidden=64 with pooling; features are the NCHW-flattened map (view(-1, fc_neurons))."""
    if dataset == 'omni':
        return NetSpec(1, 28, 28, 32, ways, 4, False, 'none')
    return NetSpec(3, 84, 84, 64, ways, 4, True, 'none')


def init_flat_params(spec, seed=42):
    """Flat fp32 parameter vector initialised like the reference constructors under
    ``torch.manual_seed(seed)`` (``vision/maml_vision.py:57,73``): per block BatchNorm2d with
    ``uniform_(weight)`` (vision_models.py:175), Conv2d default init then Xavier-uniform / zero bias
    (:186, :204-207); then the head -- Xavier/zero for MiniImagenetCNN (:103-104), ``normal_()`` weight and
    zero bias for OmniglotCNN (:47-49).  Same RNG consumption order as the reference, so the same seed
    gives the same parameters."""
    import torch
    torch.manual_seed(seed)
    parts, cin = [], spec.in_c
    for _ in range(spec.layers):
        bn = torch.nn.BatchNorm2d(spec.hidden, affine=True)
        torch.nn.init.uniform_(bn.weight)
        conv = torch.nn.Conv2d(cin, spec.hidden, (3, 3), stride=1 if spec.pool else 2, padding=1, bias=True)
        torch.nn.init.xavier_uniform_(conv.weight.data, gain=1.0)
        torch.nn.init.constant_(conv.bias.data, 0.0)
        parts += [bn.weight, bn.bias, conv.weight, conv.bias]
        cin = spec.hidden
    if spec.head != 'none':
        lin = torch.nn.Linear(spec.feat_dim(), spec.ways, bias=True)
        if spec.head == 'flatten':
            torch.nn.init.xavier_uniform_(lin.weight.data, gain=1.0)
            torch.nn.init.constant_(lin.bias.data, 0.0)
        else:
            lin.weight.data.normal_()
            lin.bias.data.mul_(0.0)
        parts += [lin.weight, lin.bias]
    return torch.cat([p.detach().reshape(-1).float() for p in parts])
