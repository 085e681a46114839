"""CPU restatement of deephar/models/action.py::build_merge_model (CVPR'18 clip model): the 2-D pose
variant (exp/pennaction/eval_penn_ar_pe_merge.py:42-62) and the 3-D one (action.py:208-297, pose_dim=3).
TEST INFRASTRUCTURE (see oracle/__init__.py).

`forward(ops, weights, x, ...)`: x (B,T,H,W,3) -> 9 action probability vectors
[p1..p4, v1..v4, m] (action.py:377-396), optionally preceded by (pose, visibility).
Weight names: the backbone keeps the ReceptionNet names (Stem/, rBlock%d/, SepConv%d/, RegMap%d/,
fReMap%d/ -- action.py:117-153 re-wires the same layers), then PoseAR/, GuidedVisAR/ and the two
top-level "heat-map weighting" SeparableConv2D layers (action.py:383-396); Keras auto-name
counters keep running across the sub-models.
"""
import numpy as np

from . import reception as R
from .reception import Ctx, Weights


def action_top(ops, x):
    """action.py:14-17."""
    return ops.softmax(ops.global_max_min_pooling(x))


def act_pred_block(c, x, num_out, last=False):
    """action.py:20-42 (include_top=False: returns the raw action heat-map)."""
    ops = c.ops
    nf = x.shape[-1]
    ident = x
    x = c.act_conv_bn(x, int(nf / 2), (1, 1))
    x = c.act_conv_bn(x, nf, (3, 3))
    x = ident + x
    ident = x
    x1 = c.act_conv_bn(x, nf, (3, 3))
    x = ops.max_min_pooling(x1, (2, 2))
    action_hm = c.act_conv(x, num_out, (3, 3))
    y = action_hm
    if not last:
        action_hm = ops.upsample2d(action_hm, (2, 2))
        action_hm = c.act_conv_bn(action_hm, nf, (3, 3))
        x = ident + x1 + action_hm
    return x, y


POSE_NET_WIDTHS = {'v1': (8, 16, 24, 56, 32), 'v2': (12, 24, 36, 112, 64)}      # action.py:55-72


def pose_model(c0, y, p, num_actions, network_version='v1'):
    """action.py:45-90 (include_top=False).  y (B,T,nj,dim), p (B,T,nj,1); 'v2' (action.py:63-72) is the wider net of
    exp/ntu/eval_ntu_ar_pe_merge.py."""
    ops = c0.ops
    c = c0.sub('PoseAR')
    w31, w33, w35, wide, squeeze = POSE_NET_WIDTHS[network_version]
    x = y * p
    a = c.conv_bn_act(x, w31, (3, 1))
    b = c.conv_bn_act(x, w33, (3, 3))
    cc = c.conv_bn_act(x, w35, (3, 5))
    x = ops.concat([a, b, cc])
    a = c.conv_bn(x, wide, (3, 3))
    b = c.conv_bn(x, squeeze, (1, 1))
    b = c.conv_bn(b, wide, (3, 3))
    x = ops.concat([a, b])
    x = ops.max_min_pooling(x, (2, 2))
    outs = []
    for i in range(4):
        x, yi = act_pred_block(c, x, num_actions, last=(i == 3))
        outs.append(yi)
    return outs


def visual_model(c0, f, num_actions):
    """action.py:93-109 (include_top=False).  f (B,T,nj,F)."""
    ops = c0.ops
    c = c0.sub('GuidedVisAR')
    x = c.conv_bn(f, 256, (1, 1))
    x = ops.maxpool2d(x, (2, 2), None, 'valid')
    outs = []
    for i in range(4):
        x, yi = act_pred_block(c, x, num_actions, last=(i == 3))
        outs.append(yi)
    return outs


def forward(ops, weight_table, x, num_actions, num_joints, num_blocks, num_context_per_joint=2,
            ksize=(5, 5), output_poses=False, weighted_merge=True, return_weights_used=False,
            pose_dim=2, depth_maps=8, pose_net_version='v1'):
    """action.py:319-400 on a ReceptionNet built as in eval_penn_ar_pe_merge.py:51-53 (pose_dim=2) or with
    dim=3, depth_maps=`depth_maps` (pose_dim=3, action.py:208-297)."""
    w = Weights(weight_table, ops)
    c = Ctx(ops, w)
    x = ops.from_numpy(x)
    B, T = x.shape[:2]
    x = x.reshape((B * T,) + tuple(x.shape[2:]))

    # ---- _get_2d_pose_estimation_from_model (action.py:112-201), TimeDistributed folded ----
    x1 = R._stem(c, x)
    xb1 = R._reception_block(c, x1, 'rBlock1', ksize)
    nfilt = xb1.shape[-1]
    num_heatmaps = (num_context_per_joint + 1) * num_joints if pose_dim == 2 else depth_maps * num_joints

    def sepconv_blk(t, i):
        return c.sub('SepConv%d' % i).separable_act_conv_bn(t, nfilt, ksize)

    def regmap(t, i):
        return c.sub('RegMap%d' % i).act_conv(t, num_heatmaps, (1, 1))

    def fremap(t, i):
        return c.sub('fReMap%d' % i).act_conv_bn(t, nfilt, (1, 1))

    # weights must be created in the reception.build order: rBlock_i, SepConv_i, RegMap_i, fReMap_i
    x2 = sepconv_blk(xb1, 1)
    x3 = fremap(regmap(x2, 1), 1)
    xx = xb1 + x2 + x3
    for i in range(2, num_blocks):
        t1 = R._reception_block(c, xx, 'rBlock%d' % i, ksize)
        t2 = sepconv_blk(t1, i)
        t3 = fremap(regmap(t2, i), i)
        xx = t1 + t2 + t3
    xx = R._reception_block(c, xx, 'rBlock%d' % num_blocks, ksize)
    xx = sepconv_blk(xx, num_blocks)
    h = regmap(xx, num_blocks)

    if pose_dim == 2:
        hs = h[..., :num_joints]
        hc = h[..., num_joints:]
        ys = R.softargmax_2d_model(ops, hs)
        yc = R.softargmax_2d_model(ops, hc)
        pc = R.joints_probability_model(ops, hc)
        y = R.context_aggregation_model(ops, ys, yc, pc, num_joints, num_context_per_joint, 0.8)
        p = R.joints_probability_model(ops, 4 * hs)                     # action.py:200
        hs_prob = ops.channel_softmax_2d(hs)                            # action.py:202-203
    else:
        # action.py:265-295: the reception 3-D head with visible = sigmoid(2 * (vxy + vz))
        n_, hh, ww, ch = h.shape
        h5 = h.reshape(n_, hh, ww, depth_maps, num_joints)
        hxy = ops.mean(h5, 3)
        hz = ops.mean(h5, (1, 2))
        y = ops.concat([R.softargmax_2d_model(ops, hxy), R.softargmax_1d_model(ops, hz)])
        p = ops.sigmoid(2 * (ops.amax(hxy, (1, 2)) + ops.amax(hz, 1)))[..., None]
        hs_prob = ops.channel_softmax_2d(hxy)

    unf = lambda t: t.reshape((B, T) + tuple(t.shape[1:]))
    y, p = unf(y), unf(p)
    outputs = []
    if output_poses:
        outputs += [y, p]

    out_pose = pose_model(c, y, p, num_actions, pose_net_version)
    f = ops.kronecker_prod(unf(hs_prob), unf(xb1))
    out_vis = visual_model(c, f, num_actions)

    for t in out_pose:
        outputs.append(action_top(ops, t))
    for t in out_vis:
        outputs.append(action_top(ops, t))

    pm, vm = out_pose[-1], out_vis[-1]
    if weighted_merge:                                              # action.py:377-393
        n = pm.shape[-1]
        pm = c.sepconv(pm, n, (1, 1))
        vm = c.sepconv(vm, n, (1, 1))
    outputs.append(action_top(ops, pm + vm))

    outputs = [ops.to_numpy(o) for o in outputs]
    if return_weights_used:
        return outputs, w.used
    return outputs
