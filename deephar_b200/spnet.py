"""B200-native drop-in for deephar/models/spnet.py (SPNet v1, TPAMI'20 multitask pose + action).

`build(cfg)` keeps the reference signature (spnet.py:355) and output ordering: pose outputs
(B,T,nj,dim+1) named dp1_pb1.. in pyramid order, then the action outputs (B,n_act)
(spnet.py:398-408); `split_model` / `get_num_predictions` follow
spnet.py:413-448.  The layer graph is the reference's, layer for layer (names included, so
`load_weights(by_name=True)` semantics carry over); model.py compiles it into fused kernels.
"""
import numpy as np

from .common import add_tensorlist, concat_tensorlist, downscaling, residual, upscaling
from .config import ModelConfig
from .graph import Graph
from .layers import (BatchNormalization, UpSampling2D, ZeroPadding2D, add, appstr,
                     channel_softmax_2d, concatenate, conv2d, depth_expectation, frames_to_clip,
                     global_max_min_pooling, keypoint_confidence, kronecker_prod, mask_multiply,
                     max_min_pooling, maxpooling2d, relu, sepconv2d, softargmax2d, softmax_lastaxis)


def prediction_branch(x, cfg, pred_activate=True, replica=None, forward_maps=True, name=None):
    """spnet.py:24-48."""
    num_pred = cfg.num_joints
    num_features = x.channels

    x = relu(x, name=appstr(name, '_act1'))
    pred_maps = conv2d(x, num_pred, (1, 1), name=appstr(name, '_conv1'))

    if replica:
        replica = conv2d(x, num_pred, (1, 1), name=appstr(name, '_conv1_replica'))

    if forward_maps:
        x = conv2d(x, num_pred, (1, 1), name=appstr(name, '_fw_maps'))
        x = concatenate([x, pred_maps])
    else:
        x = pred_maps

    if pred_activate:
        x = relu(x, name=appstr(name, '_act2'))
    x = conv2d(x, num_features, (1, 1), name=appstr(name, '_conv2'))

    return x, pred_maps, replica


def action_prediction_early_fusion(xa, p, c, af, cfg, name=None):
    """spnet.py:51-148.  p, c, af are clip tensors (T, nj, .)."""
    num_actions = cfg.num_actions
    num_features = max(cfg.num_pose_features, cfg.num_visual_features)
    num_pose_features = cfg.num_pose_features
    num_visual_features = cfg.num_visual_features
    shortname = name[0:7] if name is not None else None
    action = []

    def _individual_action_prediction(hlist, name=None):
        for i in range(len(hlist)):
            x = global_max_min_pooling(hlist[i])
            x = softmax_lastaxis(x, name=appstr(name, '%d' % i))
            action.append(x)

    def _prediction(x, name=None, shortname=None):
        num_features = x.channels
        ident = x
        x = BatchNormalization(x, name=appstr(name, '_bn1'))
        x = relu(x, name=appstr(name, '_act1'))
        x1 = conv2d(x, num_features, (3, 3), name=appstr(name, '_conv1'))

        x = max_min_pooling(x1, (2, 2))
        x = BatchNormalization(x, name=appstr(name, '_bn2'))
        x = relu(x, name=appstr(name, '_act2'))
        hlist = []
        for i in range(len(num_actions)):
            nact = num_actions[i]
            h = conv2d(x, nact, (3, 3), name=appstr(name, '_conv2h%d' % i))
            hlist.append(h)

        _individual_action_prediction(hlist, name=shortname)
        h = concat_tensorlist(hlist)

        x = UpSampling2D(h, (2, 2))
        x = relu(x, name=appstr(name, '_act3'))
        x = conv2d(x, num_features, (3, 3), name=appstr(name, '_conv3'))
        x = add([ident, x1, x])
        return x

    # padding strategy (spnet.py:98-107)
    num_frames, num_joints = p.shape[0], p.shape[1]
    time_stride = 2 if num_frames >= 16 else 1
    get_pad = lambda div, n: int(div * np.ceil(n / div) - n)
    joints_pad = get_pad(4, num_joints)
    frames_pad = get_pad(2 * time_stride, num_frames)
    top_pad = frames_pad // 2
    bottom_pad = (frames_pad + 1) // 2
    left_pad = joints_pad // 2
    right_pad = (joints_pad + 1) // 2

    # pose features
    x = mask_multiply(p, c)
    a = conv2d(x, num_pose_features // 16, (3, 1), name=appstr(name, '_p_conv0a'))
    b = conv2d(x, num_pose_features // 8, (3, 3), name=appstr(name, '_p_conv0b'))
    cc = conv2d(x, num_pose_features // 4, (3, 5), name=appstr(name, '_p_conv0c'))
    x = concatenate([a, b, cc])

    x = residual(x, (3, 3), out_size=num_pose_features, convtype='normal', features_div=2,
                 name=appstr(name, '_r1'))

    if top_pad + bottom_pad + left_pad + right_pad > 0:
        x = ZeroPadding2D(x, ((top_pad, bottom_pad), (left_pad, right_pad)))
    x1 = maxpooling2d(x, (2, 2), strides=(time_stride, 2))

    # appearance features
    x = conv2d(af, num_visual_features, (1, 1), name=appstr(name, '_v_conv0'))
    if top_pad + bottom_pad + left_pad + right_pad > 0:
        x = ZeroPadding2D(x, ((top_pad, bottom_pad), (left_pad, right_pad)))
    x2 = maxpooling2d(x, (2, 2), strides=(time_stride, 2))

    # feature fusion
    fusion = [x1, x2]
    if xa is not None:
        fusion.append(xa)
    x = concat_tensorlist(fusion)
    x = residual(x, (3, 3), out_size=num_features, convtype='normal', features_div=4,
                 name=appstr(name, '_r2'))

    xa = _prediction(x, name=appstr(name, '_pred'), shortname=appstr(shortname, '_a'))
    return action, xa


def prediction_block(xp, xa, zp, outlist, cfg, do_action, name=None):
    """spnet.py:151-248 (the dbg_decoupled_* debug outputs are not built)."""
    g = xp.g
    dim = cfg.dim
    kernel_size = cfg.kernel_size
    xmin, ymin = cfg.xmin, cfg.ymin
    sam_alpha = cfg.sam_alpha
    num_features = xp.channels
    replica = cfg.pose_replica and do_action

    xp = residual(xp, kernel_size, name=appstr(name, '_r1'))
    reinject = [xp]

    xp = BatchNormalization(xp, name=appstr(name, '_bn1'))
    xp = relu(xp, name=appstr(name, '_act1'))
    xp = sepconv2d(xp, num_features, kernel_size, name=appstr(name, '_conv1'))
    reinject.append(xp)

    xp = BatchNormalization(xp, name=appstr(name, '_bn2'))

    # 2D pose estimation
    x1, org_h, rep_h = prediction_branch(xp, cfg, pred_activate=True, replica=replica,
                                         name=appstr(name, '_heatmaps'))
    reinject.append(x1)

    h = channel_softmax_2d(org_h, alpha=sam_alpha, name=appstr(name, '_probmaps'))
    p = softargmax2d(h, limits=(xmin, ymin, 1 - xmin, 1 - ymin), name=appstr(name, '_xy'))
    c = keypoint_confidence(h, name=appstr(name, '_vis'))

    # depth estimation
    if dim == 3:
        x1, org_d, rep_d = prediction_branch(xp, cfg, pred_activate=False, replica=replica,
                                             forward_maps=False, name=appstr(name, '_depthmaps'))
        reinject.append(x1)
        z = depth_expectation(org_d, h)
        p = concatenate([p, z], name=appstr(name, '_xyz'))

    # visual features (for action only)
    action = []
    if do_action:
        g.act_cnt = getattr(g, 'act_cnt', 0) + 1       # global `act_cnt` in the reference (spnet.py:210-214)
        act_name = 'act%d' % g.act_cnt

        act_h = rep_h if replica else org_h
        act_h = channel_softmax_2d(act_h, alpha=sam_alpha, name=appstr(act_name, '_probmaps2'))
        act_p = softargmax2d(act_h, limits=(xmin, ymin, 1 - xmin, 1 - ymin), name=appstr(act_name, '_xy2'))
        act_c = keypoint_confidence(act_h, name=appstr(act_name, '_vis2'))

        if dim == 3:
            act_d = rep_d if replica else org_d
            act_z = depth_expectation(act_d, act_h)
            act_p = concatenate([act_p, act_z], name=appstr(act_name, '_xyz2'))

        af = kronecker_prod(act_h, zp, name=appstr(act_name, '_kron'))

        action, xa = action_prediction_early_fusion(xa, frames_to_clip(act_p), frames_to_clip(act_c),
                                                    frames_to_clip(af), cfg,
                                                    name=appstr(act_name, '_action'))

    xp = add_tensorlist(reinject)
    outlist[0].append(concatenate([p, c], name=name))
    if do_action:
        outlist[1] += action

    return xp, xa


def downscaling_pyramid(lp, la, lzp, outlist, cfg, do_action, name=None):
    """spnet.py:251-281."""
    assert len(lp) == len(la), 'Pose and action must have the same number of levels!'
    xp = lp[0]
    xa = la[0]
    if lzp[0] is None:
        lzp[0] = xp

    for i in range(1, len(lp)):
        num_features = xp.channels + cfg.growth

        xp = downscaling(xp, cfg, out_size=num_features, name=appstr(name, '_du%d' % i))

        if lzp[i] is None:
            lzp[i] = xp

        if lp[i] is not None:
            xp = add([xp, lp[i]])

        if xa is not None and do_action:
            xa = residual(xa, (3, 3), name=appstr(name, '_du%d_action_r0' % i))
            if la[i] is not None:
                xa = add([xa, la[i]])

        xp, xa = prediction_block(xp, xa, lzp[i], outlist, cfg, do_action,
                                  name=appstr(name, '_pb%d' % i))

        lp[i] = xp  # lateral pose connection
        la[i] = xa  # lateral action connection


def upscaling_pyramid(lp, la, lzp, outlist, cfg, do_action, name=None):
    """spnet.py:284-314."""
    assert len(lp) == len(la), 'Pose and action must have the same number of levels!'
    xp = lp[-1]
    xa = la[-1]
    if lzp[0] is None:
        lzp[0] = xp

    for i in range(len(lp) - 1)[::-1]:
        num_features = xp.channels - cfg.growth

        xp = upscaling(xp, cfg, out_size=num_features, name=appstr(name, '_uu%d' % i))

        if lzp[i] is None:
            lzp[i] = xp

        if lp[i] is not None:
            xp = add([xp, lp[i]])

        if xa is not None and do_action:
            xa = residual(xa, (3, 3), name=appstr(name, '_uu%d_action_r0' % i))
            if la[i] is not None:
                xa = add([xa, la[i]])

        xp, xa = prediction_block(xp, xa, lzp[i], outlist, cfg, do_action,
                                  name=appstr(name, '_pb%d' % i))

        lp[i] = xp  # lateral pose connection
        la[i] = xa  # lateral action connection


def entry_flow(x, cfg):
    """spnet.py:317-352."""
    growth = cfg.growth
    image_div = cfg.image_div
    downsampling_type = cfg.downsampling_type

    assert (image_div & (image_div - 1) == 0) and image_div >= 4, \
        'Invalid image_div ({}).'.format(image_div)
    assert downsampling_type in ['maxpooling', 'conv'], \
        'Invalid downsampling_type ({}).'.format(downsampling_type)
    if downsampling_type != 'maxpooling':
        raise NotImplementedError("downsampling_type='conv' is not used by the reference scripts")

    x = conv2d(x, 64, (7, 7), strides=(2, 2), name='conv1')
    x = residual(x, (3, 3), out_size=growth, convtype='normal', name='res0')
    x = maxpooling2d(x, (3, 3), strides=(2, 2))

    x = residual(x, (3, 3), out_size=2 * growth, convtype='normal', name='res1')
    x = residual(x, (3, 3), out_size=2 * growth, convtype='normal', name='res2')

    num_features = 2 * growth
    res_cnt = 2
    div_factor = 4

    while div_factor < image_div:
        num_features += growth
        x = maxpooling2d(x, (2, 2), strides=(2, 2))
        x = residual(x, (3, 3), out_size=num_features, convtype='normal', name='res%d' % (res_cnt + 1))
        x = residual(x, (3, 3), out_size=num_features, convtype='normal', name='res%d' % (res_cnt + 2))
        res_cnt += 2
        div_factor *= 2

    return x


def build(cfg, stop_grad_stem=False):
    """Sequential Pyramid Networks for 3D human pose estimation and action recognition
    (spnet.py:355-410).  `stop_grad_stem` only matters for training and is ignored."""
    from .model import Model

    assert type(cfg) == ModelConfig, 'type(cfg) ({}) is not ModelConfig'.format(type(cfg))
    input_shape = cfg.input_shape
    assert len(input_shape) in [3, 4], 'Invalid input_shape ({})'.format(input_shape)
    if cfg.dbg_decoupled_pose or cfg.dbg_decoupled_h:
        raise NotImplementedError('dbg_decoupled_* debug outputs are not built (SURVEY App. C.7)')

    g = Graph('SPNet')
    outlist = []  # Holds [[poses], [action1], [actions2], ...]
    for i in range(len(cfg.num_actions) + 1):
        outlist.append([])

    if len(input_shape) == 3:
        num_rows, num_cols, _ = input_shape
        inp = g.input(tuple(input_shape))
    else:
        num_frames, num_rows, num_cols, _ = input_shape
        g.frames_per_clip = int(num_frames)
        inp = g.input(tuple(input_shape[1:]))

    cfg.xmin = 1 / (2 * num_cols)
    cfg.ymin = 1 / (2 * num_rows)

    x = entry_flow(inp, cfg)

    lp, la, lzp = [], [], []
    for i in range(cfg.num_levels):
        lp.append(None)
        la.append(None)
        lzp.append(None)

    lp[0] = x
    for pyr in range(cfg.num_pyramids):
        do_action = (pyr + 1) in cfg.action_pyramids
        if do_action and len(input_shape) != 4:
            raise ValueError('action recognition needs clip input (T,H,W,3): kronecker_prod is only '
                             'defined for clip tensors in the reference (layers.py:478-508)')
        if pyr % 2 == 0:  # Even pyramids (0, 2, ...)
            downscaling_pyramid(lp, la, lzp, outlist, cfg, do_action, name='dp%d' % (pyr + 1))
        else:  # Odd pyramids (1, 3, ...)
            upscaling_pyramid(lp, la, lzp, outlist, cfg, do_action, name='up%d' % (pyr + 1))

    outputs = []
    for o in outlist:
        outputs += o
    g.outputs = outputs
    g.num_pose_outputs = len(outlist[0])

    key = 'spnet_j%d_d%d_p%d_a%s_r%d_f%d' % (cfg.num_joints, cfg.dim, cfg.num_pyramids,
                                              '-'.join(str(a) for a in cfg.action_pyramids),
                                              int(bool(cfg.pose_replica)), cfg.num_pose_features)
    m = Model(g, calib_key=key, name='SPNet')
    m.cfg = cfg
    return m


def get_num_predictions(num_pyramids, num_levels):
    """spnet.py:413-414."""
    return num_pyramids * (num_levels - 1)


def split_model(full_model, cfg, interlaced=False, model_names=[None, None]):
    """spnet.py:417-448: [pose model, action model] sharing the same compiled network."""
    num_pose_pred = get_num_predictions(cfg.num_pyramids, cfg.num_levels)
    num_act_pred = get_num_predictions(len(cfg.action_pyramids), cfg.num_levels)
    assert len(full_model.outputs) == num_pose_pred + len(cfg.num_actions) * num_act_pred, \
        'The given model and config are not compatible!'
    assert num_act_pred > 0, 'You are trying to split a "pose only" model.'
    if interlaced:
        raise NotImplementedError('interlaced=True is only used by the training scripts')
    modelp = full_model.output_subset(range(0, num_pose_pred), name=model_names[0])
    modela = full_model.output_subset(range(num_pose_pred, len(full_model.outputs)), name=model_names[1])
    return [modelp, modela]
