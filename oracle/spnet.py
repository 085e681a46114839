"""CPU restatement of deephar/models/spnet.py + models/common.py (TPAMI'20 multitask SPNet).
TEST INFRASTRUCTURE (see oracle/__init__.py).

`forward(ops, weights, x, cfg)` mirrors `spnet.build(cfg)` (spnet.py:355-410) applied to a batch
x of shape (B,T,H,W,3) (clip models, TimeDistributed folds B*T; layers.py:66-104) or (N,H,W,3).
Returns the output list in the reference order: pose outputs (B,T,nj,dim+1) then action
outputs (B,n_act) (spnet.py:404-408).

All layers are named explicitly in the reference, so weight names are "<layer>/<weight>".
BatchNormalization here is the Keras default (scale=True: gamma, beta, mean, var).
"""
import numpy as np

from .reception import Ctx, Weights


class ModelConfig(object):
    """deephar/config.py:150-192 (same constructor)."""

    def __init__(self, input_shape, poselayout, num_actions=[], num_pyramids=8, action_pyramids=[1, 2],
                 num_levels=4, kernel_size=(5, 5), growth=96, image_div=8, predict_rootz=False,
                 downsampling_type='maxpooling', pose_replica=False, num_pose_features=128,
                 num_visual_features=128, sam_alpha=1, dbg_decoupled_pose=False, dbg_decoupled_h=False):
        self.input_shape = input_shape
        self.num_joints = poselayout.num_joints
        self.dim = poselayout.dim
        assert type(num_actions) == list, 'num_actions should be a list'
        self.num_actions = num_actions
        self.num_pyramids = num_pyramids
        self.action_pyramids = action_pyramids
        self.num_levels = num_levels
        self.kernel_size = kernel_size
        self.growth = growth
        self.image_div = image_div
        self.predict_rootz = predict_rootz
        self.downsampling_type = downsampling_type
        self.pose_replica = pose_replica
        self.num_pose_features = num_pose_features
        self.num_visual_features = num_visual_features
        self.sam_alpha = sam_alpha
        self.dbg_decoupled_pose = dbg_decoupled_pose
        self.dbg_decoupled_h = dbg_decoupled_h


class pa16j2d(object):      # utils/pose.py:127-128
    num_joints, dim = 16, 2


class pa17j3d(object):      # utils/pose.py:136-137
    num_joints, dim = 17, 3


class pa20j3d(object):
    num_joints, dim = 20, 3


class _State(object):
    def __init__(self):
        self.act_cnt = 0


def residual_unit(c, x, kernel_size, strides=(1, 1), out_size=None, convtype='depthwise',
                  shortcut_act=True, features_div=2, name=None):
    """models/common.py:25-67."""
    ops = c.ops
    num_filters = x.shape[-1]
    if out_size is None:
        out_size = num_filters
    skip_conv = (num_filters != out_size) or (strides != (1, 1))
    if skip_conv:
        x = c.bn(x, True, name + '_bn1')
    shortcut = x
    if skip_conv:
        if shortcut_act:
            shortcut = ops.relu(shortcut)
        shortcut = c.conv(shortcut, out_size, (1, 1), strides=strides, name=name + '_shortcut_conv')
    if not skip_conv:
        x = c.bn(x, True, name + '_bn1')
    x = ops.relu(x)
    if convtype == 'depthwise':
        x = c.sepconv(x, out_size, kernel_size, strides=strides, name=name + '_conv1')
    else:
        x = c.conv(x, int(out_size / features_div), (1, 1), name=name + '_conv1')
        x = c.bn(x, True, name + '_bn2')
        x = ops.relu(x)
        x = c.conv(x, out_size, kernel_size, strides=strides, name=name + '_conv2')
    return shortcut + x


def prediction_branch(c, x, cfg, pred_activate=True, replica=None, forward_maps=True, name=None):
    """spnet.py:24-48."""
    ops = c.ops
    num_pred = cfg.num_joints
    num_features = x.shape[-1]
    x = ops.relu(x)
    pred_maps = c.conv(x, num_pred, (1, 1), name=name + '_conv1')
    if replica:
        replica = c.conv(x, num_pred, (1, 1), name=name + '_conv1_replica')
    if forward_maps:
        x = c.conv(x, num_pred, (1, 1), name=name + '_fw_maps')
        x = ops.concat([x, pred_maps])
    else:
        x = pred_maps
    if pred_activate:
        x = ops.relu(x)
    x = c.conv(x, num_features, (1, 1), name=name + '_conv2')
    return x, pred_maps, replica


def _heads(ops, hmap, dmap, alpha, B, T):
    """softmax -> soft-argmax -> confidence [-> depth expectation] (spnet.py:178-205)."""
    h = ops.channel_softmax_2d(hmap, alpha)
    p = ops.softargmax2d(h)
    cf = ops.keypoint_confidence(h)
    if dmap is not None:
        d = ops.sigmoid(dmap)
        z = ops.asum(d * h, (1, 2))[..., None]
        p = ops.concat([p, z])
    return h, p, cf


def action_prediction_early_fusion(c, xa, p, cf, af, cfg, name):
    """spnet.py:51-148.  p (B,T,nj,dim), cf (B,T,nj,1), af (B,T,nj,F); xa (B,T',J',nf) or None."""
    ops = c.ops
    num_actions = cfg.num_actions
    num_features = max(cfg.num_pose_features, cfg.num_visual_features)
    npf, nvf = cfg.num_pose_features, cfg.num_visual_features
    shortname = name[0:7]
    action = []

    num_frames, num_joints = p.shape[1:3]
    time_stride = 2 if num_frames >= 16 else 1
    get_pad = lambda div, n: int(div * np.ceil(n / div) - n)
    joints_pad = get_pad(4, num_joints)
    frames_pad = get_pad(2 * time_stride, num_frames)
    top_pad, bottom_pad = frames_pad // 2, (frames_pad + 1) // 2
    left_pad, right_pad = joints_pad // 2, (joints_pad + 1) // 2
    pads = ((top_pad, bottom_pad), (left_pad, right_pad))
    has_pad = top_pad + bottom_pad + left_pad + right_pad > 0

    x = p * cf                                          # mask = tile(c) ; x = p * mask
    a = c.conv(x, npf // 16, (3, 1), name=name + '_p_conv0a')
    b = c.conv(x, npf // 8, (3, 3), name=name + '_p_conv0b')
    cc = c.conv(x, npf // 4, (3, 5), name=name + '_p_conv0c')
    x = ops.concat([a, b, cc])
    x = residual_unit(c, x, (3, 3), out_size=npf, convtype='normal', features_div=2, name=name + '_r1')
    if has_pad:
        x = ops.zeropad2d(x, pads)
    x1 = ops.maxpool2d(x, (2, 2), (time_stride, 2), 'same')

    x = c.conv(af, nvf, (1, 1), name=name + '_v_conv0')
    if has_pad:
        x = ops.zeropad2d(x, pads)
    x2 = ops.maxpool2d(x, (2, 2), (time_stride, 2), 'same')

    fusion = [x1, x2]
    if xa is not None:
        fusion.append(xa)
    x = ops.concat(fusion) if len(fusion) > 1 else fusion[0]
    x = residual_unit(c, x, (3, 3), out_size=num_features, convtype='normal', features_div=4,
                      name=name + '_r2')

    # _prediction (spnet.py:71-96)
    pname = name + '_pred'
    nf = x.shape[-1]
    ident = x
    x = c.bn(x, True, pname + '_bn1')
    x = ops.relu(x)
    x1p = c.conv(x, nf, (3, 3), name=pname + '_conv1')
    x = ops.max_min_pooling(x1p, (2, 2))
    x = c.bn(x, True, pname + '_bn2')
    x = ops.relu(x)
    hlist = []
    for i in range(len(num_actions)):
        hlist.append(c.conv(x, num_actions[i], (3, 3), name=pname + '_conv2h%d' % i))
    for h in hlist:
        action.append(ops.softmax(ops.global_max_min_pooling(h)))
    h = ops.concat(hlist) if len(hlist) > 1 else hlist[0]
    x = ops.upsample2d(h, (2, 2))
    x = ops.relu(x)
    x = c.conv(x, nf, (3, 3), name=pname + '_conv3')
    xa = ident + x1p + x
    return action, xa


def prediction_block(c, st, xp, xa, zp, outlist, cfg, do_action, name, B, T):
    """spnet.py:151-248.  xp/zp are folded (B*T,H,W,C); xa is (B,T',J',nf)."""
    ops = c.ops
    dim = cfg.dim
    ks = cfg.kernel_size
    alpha = cfg.sam_alpha
    num_features = xp.shape[-1]
    replica = cfg.pose_replica and do_action

    xp = residual_unit(c, xp, ks, name=name + '_r1')
    reinject = [xp]
    xp = c.bn(xp, True, name + '_bn1')
    xp = ops.relu(xp)
    xp = c.sepconv(xp, num_features, ks, name=name + '_conv1')
    reinject.append(xp)
    xp = c.bn(xp, True, name + '_bn2')

    x1, org_h, rep_h = prediction_branch(c, xp, cfg, pred_activate=True, replica=replica,
                                         name=name + '_heatmaps')
    reinject.append(x1)
    org_d = rep_d = None
    if dim == 3:
        x1, org_d, rep_d = prediction_branch(c, xp, cfg, pred_activate=False, replica=replica,
                                             forward_maps=False, name=name + '_depthmaps')
        reinject.append(x1)
    h, p, cf = _heads(ops, org_h, org_d, alpha, B, T)

    action = []
    if do_action:
        st.act_cnt += 1
        act_name = 'act%d' % st.act_cnt
        act_hm = rep_h if replica else org_h
        act_dm = (rep_d if replica else org_d) if dim == 3 else None
        act_h, act_p, act_c = _heads(ops, act_hm, act_dm, alpha, B, T)
        unf = lambda t: t.reshape((B, T) + tuple(t.shape[1:]))
        af = ops.kronecker_prod(unf(act_h), unf(zp))
        action, xa = action_prediction_early_fusion(c, xa, unf(act_p), unf(act_c), af, cfg,
                                                    name=act_name + '_action')

    xs = reinject[0]
    for t in reinject[1:]:
        xs = xs + t
    pc = ops.concat([p, cf])
    outlist[0].append(pc.reshape((B, T) + tuple(pc.shape[1:])) if T > 1 or c.clip else pc)
    if do_action:
        outlist[1] += action
    return xs, xa


def entry_flow(c, x, cfg):
    """spnet.py:317-352."""
    ops = c.ops
    growth, image_div = cfg.growth, cfg.image_div
    assert (image_div & (image_div - 1) == 0) and image_div >= 4
    assert cfg.downsampling_type == 'maxpooling', 'only maxpooling is used by the shipped scripts'
    x = c.conv(x, 64, (7, 7), strides=(2, 2), name='conv1')
    x = residual_unit(c, x, (3, 3), out_size=growth, convtype='normal', name='res0')
    x = ops.maxpool2d(x, (3, 3), (2, 2), 'same')
    x = residual_unit(c, x, (3, 3), out_size=2 * growth, convtype='normal', name='res1')
    x = residual_unit(c, x, (3, 3), out_size=2 * growth, convtype='normal', name='res2')
    num_features = 2 * growth
    res_cnt = 2
    div_factor = 4
    while div_factor < image_div:
        num_features += growth
        x = ops.maxpool2d(x, (2, 2), (2, 2), 'same')
        x = residual_unit(c, x, (3, 3), out_size=num_features, convtype='normal', name='res%d' % (res_cnt + 1))
        x = residual_unit(c, x, (3, 3), out_size=num_features, convtype='normal', name='res%d' % (res_cnt + 2))
        res_cnt += 2
        div_factor *= 2
    return x


def forward(ops, weight_table, x, cfg, return_weights_used=False):
    """spnet.py:355-410."""
    input_shape = cfg.input_shape
    assert len(input_shape) in [3, 4]
    w = Weights(weight_table, ops)
    c = Ctx(ops, w)
    c.clip = len(input_shape) == 4
    x = ops.from_numpy(x)
    if c.clip:
        B, T = x.shape[:2]
        x = x.reshape((B * T,) + tuple(x.shape[2:]))
    else:
        B, T = x.shape[0], 1
    st = _State()
    outlist = [[] for _ in range(len(cfg.num_actions) + 1)]

    x = entry_flow(c, x, cfg)
    lp = [None] * cfg.num_levels
    la = [None] * cfg.num_levels
    lzp = [None] * cfg.num_levels
    lp[0] = x
    ks = cfg.kernel_size
    for pyr in range(cfg.num_pyramids):
        do_action = (pyr + 1) in cfg.action_pyramids
        if pyr % 2 == 0:                                   # downscaling_pyramid (spnet.py:251-281)
            name = 'dp%d' % (pyr + 1)
            xp, xa = lp[0], la[0]
            if lzp[0] is None:
                lzp[0] = xp
            for i in range(1, len(lp)):
                nfeat = xp.shape[-1] + cfg.growth
                xp = ops.maxpool2d(xp, (2, 2), (2, 2), 'same')                   # common.py:70-86
                xp = residual_unit(c, xp, ks, out_size=nfeat, name=name + '_du%d' % i + '_r0')
                if lzp[i] is None:
                    lzp[i] = xp
                if lp[i] is not None:
                    xp = xp + lp[i]
                if xa is not None and do_action:
                    xa = residual_unit(c, xa, (3, 3), name=name + '_du%d_action_r0' % i)
                    if la[i] is not None:
                        xa = xa + la[i]
                xp, xa = prediction_block(c, st, xp, xa, lzp[i], outlist, cfg, do_action,
                                          name + '_pb%d' % i, B, T)
                lp[i], la[i] = xp, xa
        else:                                              # upscaling_pyramid (spnet.py:284-314)
            name = 'up%d' % (pyr + 1)
            xp, xa = lp[-1], la[-1]
            if lzp[0] is None:
                lzp[0] = xp
            for i in range(len(lp) - 1)[::-1]:
                nfeat = xp.shape[-1] - cfg.growth
                xp = ops.upsample2d(xp, (2, 2))                                  # common.py:89-108
                xp = residual_unit(c, xp, ks, out_size=nfeat, name=name + '_uu%d' % i + '_r0')
                if lzp[i] is None:
                    lzp[i] = xp
                if lp[i] is not None:
                    xp = xp + lp[i]
                if xa is not None and do_action:
                    xa = residual_unit(c, xa, (3, 3), name=name + '_uu%d_action_r0' % i)
                    if la[i] is not None:
                        xa = xa + la[i]
                xp, xa = prediction_block(c, st, xp, xa, lzp[i], outlist, cfg, do_action,
                                          name + '_pb%d' % i, B, T)
                lp[i], la[i] = xp, xa

    outputs = []
    for o in outlist:
        outputs += o
    outputs = [ops.to_numpy(o) for o in outputs]
    if return_weights_used:
        return outputs, w.used
    return outputs


def get_num_predictions(num_pyramids, num_levels):
    """spnet.py:413-414."""
    return num_pyramids * (num_levels - 1)
