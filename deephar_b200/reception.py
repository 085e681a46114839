"""B200-native drop-in for deephar/models/reception.py (CVPR'18 ReceptionNet).

`build(...)` keeps the reference signature (reception.py:225-234) and returns a
`deephar_b200.model.Model` whose `.predict` / `.outputs` / `.input_shape` /
`.load_weights` follow the keras.Model protocol the reference evaluators use
(exp/common/mpii_tools.py:63-90, h36m_tools.py:12-50).  The layer graph recorded here is
the reference's, layer for layer; model.py compiles it into fused sm_100a kernels.
"""
from .graph import Graph
from .layers import (MaxPooling2D, UpSampling2D, act_conv, act_conv_bn, add, concatenate,
                     conv_bn, conv_bn_act, separable_act_conv_bn)


def _sepconv_residual(x, out_size, name, kernel_size=(3, 3)):
    """reception.py:43-59."""
    shortcut_name = name + '_shortcut'
    reduce_name = name + '_reduce'

    num_filters = x.channels
    if num_filters == out_size:
        ident = x
    else:
        ident = act_conv_bn(x, out_size, (1, 1), name=shortcut_name)

    if out_size < num_filters:
        x = act_conv_bn(x, out_size, (1, 1), name=reduce_name)

    x = separable_act_conv_bn(x, out_size, kernel_size, name=name)
    x = add([ident, x])
    return x


def _stem(inp, old_model=False):
    """reception.py:61-98."""
    assert not old_model, 'old_model=True is not used by any shipped script and is not supported'
    g = inp.g
    with g.scope('Stem'):
        x = conv_bn_act(inp, 32, (3, 3), strides=(2, 2))
        x = conv_bn_act(x, 32, (3, 3))
        x = conv_bn_act(x, 64, (3, 3))

        a = conv_bn_act(x, 96, (3, 3), strides=(2, 2))
        b = MaxPooling2D(x, (3, 3), strides=(2, 2), padding='same')
        x = concatenate([a, b])

        a = conv_bn_act(x, 64, (1, 1))
        a = conv_bn(a, 96, (3, 3))
        b = conv_bn_act(x, 64, (1, 1))
        b = conv_bn_act(b, 64, (5, 1))
        b = conv_bn_act(b, 64, (1, 5))
        b = conv_bn(b, 96, (3, 3))
        x = concatenate([a, b])

        a = act_conv_bn(x, 192, (3, 3), strides=(2, 2))
        b = MaxPooling2D(x, (2, 2), strides=(2, 2))
        x = concatenate([a, b])

        x = _sepconv_residual(x, 3 * 192, name='sepconv1')
    return x


def build_reception_block(inp, name, ksize=(3, 3)):
    """reception.py:101-131."""
    size = inp.channels
    with inp.g.scope(name):
        xi = inp
        a = _sepconv_residual(xi, size, name='sepconv_l1', kernel_size=ksize)

        low1 = MaxPooling2D(xi, (2, 2))
        low1 = act_conv_bn(low1, int(size / 2), (1, 1))
        low1 = _sepconv_residual(low1, int(size / 2), name='sepconv_l2_1', kernel_size=ksize)
        b = _sepconv_residual(low1, int(size / 2), name='sepconv_l2_2', kernel_size=ksize)

        c = MaxPooling2D(low1, (2, 2))
        c = _sepconv_residual(c, int(size / 2), name='sepconv_l3_1', kernel_size=ksize)
        c = _sepconv_residual(c, int(size / 2), name='sepconv_l3_2', kernel_size=ksize)
        c = _sepconv_residual(c, int(size / 2), name='sepconv_l3_3', kernel_size=ksize)
        c = UpSampling2D(c, (2, 2))

        b = add([b, c])
        b = _sepconv_residual(b, size, name='sepconv_l2_3', kernel_size=ksize)
        b = UpSampling2D(b, (2, 2))
        x = add([a, b])
    return x


def build_sconv_block(inp, name=None, ksize=(3, 3)):
    """reception.py:134-142."""
    with inp.g.scope(name):
        return separable_act_conv_bn(inp, inp.channels, ksize)


def build_regmap_block(inp, num_maps, name=None):
    """reception.py:145-153."""
    with inp.g.scope(name):
        return act_conv(inp, num_maps, (1, 1))


def build_fremap_block(inp, num_filters, name=None):
    """reception.py:156-164."""
    with inp.g.scope(name):
        return act_conv_bn(inp, num_filters, (1, 1))


def pose_regression_2d_context(h, num_joints, num_context_per_joint, alpha):
    """reception.py:167-182 with sSAM/cSAM (blocks.py:306-325), sjProb/cjProb on the RAW maps
    (blocks.py:328-343) and Agg (blocks.py:217-285) -- one parameter-free graph op, one kernel."""
    pose, visible = h.g.op('pose_regression_2d_context', [h],
                           [(1, num_joints, 2), (1, num_joints, 1)],
                           {'num_joints': num_joints, 'num_context': num_context_per_joint,
                            'alpha': float(alpha)})
    from .layers import channel_slice
    hs = channel_slice(h, 0, num_joints)
    return pose, visible, hs


def pose_regression_2d(h):
    """reception.py:185-190 (unreachable from build(): num_context_per_joint defaults to 2)."""
    c = h.channels
    pose, visible = h.g.op('pose_regression_2d', [h], [(1, c, 2), (1, c, 1)], {})
    return pose, visible, h


def pose_regression_3d(h, num_joints, depth_maps):
    """reception.py:193-222 (zSAM: blocks.py:288-303)."""
    assert h.channels == depth_maps * num_joints
    pose, visible = h.g.op('pose_regression_3d', [h], [(1, num_joints, 3), (1, num_joints, 1)],
                           {'num_joints': num_joints, 'depth_maps': depth_maps})
    return pose, visible, None


def build(input_shape, num_joints, dim,
          num_context_per_joint=None,
          alpha=0.8,
          num_blocks=4,
          depth_maps=16,
          ksize=(3, 3),
          export_heatmaps=False,
          export_vfeat_block=None,
          old_model=False,
          concat_pose_confidence=True):
    """reception.py:225-319."""
    from .model import Model

    if dim == 2:
        if num_context_per_joint is None:
            num_context_per_joint = 2
        num_heatmaps = (num_context_per_joint + 1) * num_joints
    elif dim == 3:
        assert num_context_per_joint is None, \
            'For 3D pose estimation, contextual heat maps are not allowed.'
        num_heatmaps = depth_maps * num_joints
    else:
        raise ValueError('"dim" must be 2 or 3 and not (%d)' % dim)
    if export_heatmaps and dim == 3:
        raise NotImplementedError('export_heatmaps with dim=3 (hxy marginal) is not exported')

    g = Graph('ReceptionNet')
    inp = g.input(tuple(input_shape))
    outputs = []
    vfeat = None

    x = _stem(inp, old_model=old_model)

    for bidx in range(num_blocks):
        block_shape = x.shape
        x = build_reception_block(x, name='rBlock%d' % (bidx + 1), ksize=ksize)

        if export_vfeat_block == (bidx + 1):
            vfeat = x

        ident_map = x
        x = build_sconv_block(x, name='SepConv%d' % (bidx + 1), ksize=ksize)
        h = build_regmap_block(x, num_heatmaps, name='RegMap%d' % (bidx + 1))

        if dim == 2:
            if num_context_per_joint is not None:
                pose, visible, hm = pose_regression_2d_context(h, num_joints,
                                                               num_context_per_joint, alpha)
            else:
                pose, visible, hm = pose_regression_2d(h)
        else:
            pose, visible, hm = pose_regression_3d(h, num_joints, depth_maps)

        if concat_pose_confidence:
            outputs.append(concatenate([pose, visible]))
        else:
            outputs.append(pose)
            outputs.append(visible)

        if export_heatmaps:
            outputs.append(hm)

        if bidx < num_blocks - 1:
            h = build_fremap_block(h, block_shape[-1], name='fReMap%d' % (bidx + 1))
            x = add([ident_map, x, h])

    if vfeat is not None:
        outputs.append(vfeat)

    g.outputs = outputs
    calib_key = 'reception_j%d_d%d_c%s_k%d' % (num_joints, dim, num_context_per_joint, ksize[0])
    m = Model(g, calib_key=calib_key)
    m.build_args = dict(num_joints=num_joints, dim=dim, num_context_per_joint=num_context_per_joint,
                        num_blocks=num_blocks, ksize=tuple(ksize))
    return m
