"""Static description of the EfficientNet-B3 trunk used by the pose models.

The table is derived from the published EfficientNet scaling rule (width 1.2,
depth 1.4, channel divisor 8) applied to the seven base stages, the same rule
the reference evaluates at model construction
(reference: cosypose/models/efficientnet_utils.py:59-81 round_filters/round_repeats,
:259-264 base stages, :169 B3 coefficients; cosypose/models/efficientnet.py:136-157
stage unrolling).  TF-"same" padding is *static*, computed for a 300x300 image
(reference: efficientnet_utils.py:123-146), which makes stride-2 padding
asymmetric: k3 -> (0,1), k5 -> (1,2); stride-1 is symmetric (k-1)/2.

The same table is compiled into the CUDA engine (csrc/effnet_table.h); the test
`tests/test_spec.py` checks that the two agree.
"""
import math
from collections import namedtuple

BlockSpec = namedtuple(
    'BlockSpec', 'idx k s e cin cexp cse cout pad_lo pad_hi skip')

_BASE_STAGES = [
    # repeats, kernel, stride, expand, in, out
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
]
_WIDTH, _DEPTH, _DIVISOR, _PAD_IMAGE = 1.2, 1.4, 8, 300
SE_RATIO = 0.25
BN_EPS = 1e-3
IN_CHANNELS = 6
RENDER_H, RENDER_W = 240, 320
N_FEATURES = 1536
POSE_DIM = 9


def _round_filters(f):
    f = f * _WIDTH
    new_f = max(_DIVISOR, int(f + _DIVISOR / 2) // _DIVISOR * _DIVISOR)
    if new_f < 0.9 * f:
        new_f += _DIVISOR
    return int(new_f)


def _round_repeats(r):
    return int(math.ceil(_DEPTH * r))


def _static_same_pad(k, s):
    out = math.ceil(_PAD_IMAGE / s)
    pad = max((out - 1) * s + (k - 1) + 1 - _PAD_IMAGE, 0)
    return pad // 2, pad - pad // 2


def make_blocks():
    blocks = []
    for (r, k, s, e, i, o) in _BASE_STAGES:
        cin, cout = _round_filters(i), _round_filters(o)
        for rep in range(_round_repeats(r)):
            stride = s if rep == 0 else 1
            bcin = cin if rep == 0 else cout
            lo, hi = _static_same_pad(k, stride)
            blocks.append(BlockSpec(
                idx=len(blocks), k=k, s=stride, e=e, cin=bcin, cexp=bcin * e,
                cse=max(1, int(bcin * SE_RATIO)), cout=cout, pad_lo=lo, pad_hi=hi,
                skip=(stride == 1 and bcin == cout)))
    return blocks


BLOCKS = make_blocks()
STEM_OUT = _round_filters(32)
STEM_PAD = _static_same_pad(3, 2)
HEAD_IN = BLOCKS[-1].cout


def out_size(n, k, s, lo, hi):
    return (n + lo + hi - k) // s + 1


def activation_shapes(h=RENDER_H, w=RENDER_W):
    """[(name, H, W, C)] of every block-boundary activation, stem input first."""
    shapes = [('input', h, w, IN_CHANNELS)]
    h = out_size(h, 3, 2, *STEM_PAD)
    w = out_size(w, 3, 2, *STEM_PAD)
    shapes.append(('stem', h, w, STEM_OUT))
    for b in BLOCKS:
        h = out_size(h, b.k, b.s, b.pad_lo, b.pad_hi)
        w = out_size(w, b.k, b.s, b.pad_lo, b.pad_hi)
        shapes.append((f'block{b.idx}', h, w, b.cout))
    shapes.append(('head', h, w, N_FEATURES))
    return shapes


def state_dict_layout():
    """Ordered {name: shape} of the pose model's float tensors, in the
    reference's `state_dict` naming (reference: cosypose/models/pose.py:23,33,
    cosypose/models/efficientnet.py:46-68,129-160).  `num_batches_tracked`
    scalars are omitted (int64, ignored by the engine)."""
    out = {}

    def bn(prefix, c):
        for f in ('weight', 'bias', 'running_mean', 'running_var'):
            out[f'{prefix}.{f}'] = (c,)

    out['backbone._conv_stem.weight'] = (STEM_OUT, IN_CHANNELS, 3, 3)
    bn('backbone._bn0', STEM_OUT)
    for b in BLOCKS:
        p = f'backbone._blocks.{b.idx}'
        if b.e != 1:
            out[f'{p}._expand_conv.weight'] = (b.cexp, b.cin, 1, 1)
            bn(f'{p}._bn0', b.cexp)
        out[f'{p}._depthwise_conv.weight'] = (b.cexp, 1, b.k, b.k)
        bn(f'{p}._bn1', b.cexp)
        out[f'{p}._se_reduce.weight'] = (b.cse, b.cexp, 1, 1)
        out[f'{p}._se_reduce.bias'] = (b.cse,)
        out[f'{p}._se_expand.weight'] = (b.cexp, b.cse, 1, 1)
        out[f'{p}._se_expand.bias'] = (b.cexp,)
        out[f'{p}._project_conv.weight'] = (b.cout, b.cexp, 1, 1)
        bn(f'{p}._bn2', b.cout)
    out['backbone._conv_head.weight'] = (N_FEATURES, HEAD_IN, 1, 1)
    bn('backbone._bn1', N_FEATURES)
    out['pose_fc.weight'] = (POSE_DIM, N_FEATURES)
    out['pose_fc.bias'] = (POSE_DIM,)
    return out
