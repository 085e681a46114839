"""CPU restatement of deephar/models/reception.py (+ the head blocks of
deephar/models/blocks.py:217-343).  TEST INFRASTRUCTURE (see oracle/__init__.py).

`forward(ops, weights, x, ...)` takes the same arguments as the reference's
`reception.build(...)` (reception.py:225-234) plus the input batch, and returns the
output list in the reference's order (reception.py:300-307).

Weight names follow Keras: "<sub-model>/<layer>/<weight>", unnamed layers get the
Keras auto name "<class_snake>_<n>" with one global counter per class, advanced in
the reference's layer-creation order.
"""
import numpy as np


class Weights(object):
    """Name -> array lookup that records what was consumed (names + shapes)."""

    def __init__(self, table, ops):
        self.table = table
        self.ops = ops
        self.used = []
        self.counters = {}

    def auto(self, prefix):
        n = self.counters.get(prefix, 0) + 1
        self.counters[prefix] = n
        return '%s_%d' % (prefix, n)

    def get(self, name, shape):
        if hasattr(self.table, 'lookup'):
            a = self.table.lookup(name, shape)
        else:
            a = self.table[name]
        assert tuple(a.shape) == tuple(shape), (name, a.shape, shape)
        self.used.append((name, tuple(shape)))
        return self.ops.from_numpy(a)


class Ctx(object):
    def __init__(self, ops, weights, scope=''):
        self.ops = ops
        self.w = weights
        self.scope = scope

    def sub(self, scope):
        return Ctx(self.ops, self.w, scope)

    def wname(self, layer, wn):
        return (self.scope + '/' if self.scope else '') + layer + '/' + wn

    # --- keras layers ------------------------------------------------------
    def conv(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        name = name or self.w.auto('conv2d')
        cin = x.shape[-1]
        kname = self.wname(name, 'kernel')
        k = self.w.get(kname, (size[0], size[1], cin, filters))
        y = self.ops.conv2d(x, k, strides, padding)
        tab = self.w.table
        if hasattr(tab, 'observe_head') and kname not in tab.calib:
            from .synth import is_head_kernel
            if is_head_kernel(kname):       # calibration run only (oracle/synth.py)
                tab.observe_head(kname, self.ops.to_numpy(y))
                y = self.ops.conv2d(x, self.w.get(kname, k.shape), strides, padding)
        return y

    def sepconv(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        name = name or self.w.auto('separable_conv2d')
        cin = x.shape[-1]
        dw = self.w.get(self.wname(name, 'depthwise_kernel'), (size[0], size[1], cin, 1))
        pw = self.w.get(self.wname(name, 'pointwise_kernel'), (1, 1, cin, filters))
        return self.ops.separable_conv2d(x, dw, pw, strides, padding)

    def bn(self, x, scale, name=None):
        name = name or self.w.auto('batch_normalization')
        c = x.shape[-1]
        if hasattr(self.w.table, 'observe_bn'):   # calibration run only (oracle/synth.py)
            self.w.table.observe_bn((self.scope + '/' if self.scope else '') + name,
                                    self.ops.to_numpy(x))
        gamma = self.w.get(self.wname(name, 'gamma'), (c,)) if scale else None
        beta = self.w.get(self.wname(name, 'beta'), (c,))
        mean = self.w.get(self.wname(name, 'moving_mean'), (c,))
        var = self.w.get(self.wname(name, 'moving_variance'), (c,))
        return self.ops.batchnorm(x, gamma, beta, mean, var)

    # --- layers.py combos (layers.py:202-325); BN there has scale=False ------
    def conv_bn(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        x = self.conv(x, filters, size, strides, padding, name + '_conv' if name else None)
        return self.bn(x, False, name)

    def conv_bn_act(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        x = self.conv(x, filters, size, strides, padding, name + '_conv' if name else None)
        x = self.bn(x, False, name + '_bn' if name else None)
        return self.ops.relu(x)

    def act_conv_bn(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        x = self.ops.relu(x)
        x = self.conv(x, filters, size, strides, padding, name + '_conv' if name else None)
        return self.bn(x, False, name)

    def act_conv(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        x = self.ops.relu(x)
        return self.conv(x, filters, size, strides, padding, name)

    def separable_act_conv_bn(self, x, filters, size, strides=(1, 1), padding='same', name=None):
        x = self.ops.relu(x)
        x = self.sepconv(x, filters, size, strides, padding, name + '_conv' if name else None)
        return self.bn(x, False, name)


def _sepconv_residual(c, x, out_size, name, kernel_size=(3, 3)):
    """reception.py:43-59."""
    num_filters = x.shape[-1]
    if num_filters == out_size:
        ident = x
    else:
        ident = c.act_conv_bn(x, out_size, (1, 1), name=name + '_shortcut')
    if out_size < num_filters:
        x = c.act_conv_bn(x, out_size, (1, 1), name=name + '_reduce')
    x = c.separable_act_conv_bn(x, out_size, kernel_size, name=name)
    return ident + x


def _stem(c0, inp):
    """reception.py:61-98 (old_model=False)."""
    ops = c0.ops
    c = c0.sub('Stem')
    x = c.conv_bn_act(inp, 32, (3, 3), strides=(2, 2))
    x = c.conv_bn_act(x, 32, (3, 3))
    x = c.conv_bn_act(x, 64, (3, 3))

    a = c.conv_bn_act(x, 96, (3, 3), strides=(2, 2))
    b = ops.maxpool2d(x, (3, 3), (2, 2), 'same')
    x = ops.concat([a, b])

    a = c.conv_bn_act(x, 64, (1, 1))
    a = c.conv_bn(a, 96, (3, 3))
    b = c.conv_bn_act(x, 64, (1, 1))
    b = c.conv_bn_act(b, 64, (5, 1))
    b = c.conv_bn_act(b, 64, (1, 5))
    b = c.conv_bn(b, 96, (3, 3))
    x = ops.concat([a, b])

    a = c.act_conv_bn(x, 192, (3, 3), strides=(2, 2))
    b = ops.maxpool2d(x, (2, 2), (2, 2), 'valid')
    x = ops.concat([a, b])

    x = _sepconv_residual(c, x, 3 * 192, name='sepconv1')
    return x


def _reception_block(c0, xi, name, ksize):
    """reception.py:101-131."""
    ops = c0.ops
    c = c0.sub(name)
    size = xi.shape[-1]
    a = _sepconv_residual(c, xi, size, 'sepconv_l1', ksize)

    low1 = ops.maxpool2d(xi, (2, 2), None, 'valid')
    low1 = c.act_conv_bn(low1, int(size / 2), (1, 1))
    low1 = _sepconv_residual(c, low1, int(size / 2), 'sepconv_l2_1', ksize)
    b = _sepconv_residual(c, low1, int(size / 2), 'sepconv_l2_2', ksize)

    cc = ops.maxpool2d(low1, (2, 2), None, 'valid')
    cc = _sepconv_residual(c, cc, int(size / 2), 'sepconv_l3_1', ksize)
    cc = _sepconv_residual(c, cc, int(size / 2), 'sepconv_l3_2', ksize)
    cc = _sepconv_residual(c, cc, int(size / 2), 'sepconv_l3_3', ksize)
    cc = ops.upsample2d(cc, (2, 2))

    b = b + cc
    b = _sepconv_residual(c, b, size, 'sepconv_l2_3', ksize)
    b = ops.upsample2d(b, (2, 2))
    return a + b


# --- parameter-free heads (blocks.py:217-343) --------------------------------
def softargmax_2d_model(ops, h):
    """blocks.py:306-325 (rho=0): channel softmax then the two grid convolutions."""
    return ops.softargmax2d(ops.channel_softmax_2d(h))


def joints_probability_model(ops, h):
    """blocks.py:328-343: applied to whatever it is given (RAW maps in reception)."""
    return ops.keypoint_confidence(h)


def softargmax_1d_model(ops, hz):
    """blocks.py:288-303."""
    return ops.lin_interpolation_1d(ops.channel_softmax_1d(hz))


def context_aggregation_model(ops, ys, yc, pc, num_joints, num_context, alpha):
    """blocks.py:217-285 (num_frames=1).  ys (N,nj,2), yc (N,nj*nc,2), pc (N,nj*nc,1)."""
    n = ys.shape[0]

    def ctx_sum(v):     # fixed Dense: w[j*nc:(j+1)*nc, j] = 1   (blocks.py:227-233)
        return ops.asum(v.reshape(n, num_joints, num_context, 1), 2)

    xi = yc[:, :, 0:1]
    yi = yc[:, :, 1:2]
    pc_sum = ctx_sum(pc)
    pxi_div = ctx_sum(xi * pc) / pc_sum
    pyi_div = ctx_sum(yi * pc) / pc_sum
    yc_div = ops.concat([pxi_div, pyi_div])
    return alpha * ys + (1 - alpha) * yc_div


def pose_regression_2d_context(ops, h, num_joints, num_context, alpha, debug=None):
    """reception.py:167-182.  `debug` (dict) receives the conditioning of the context division:
    the reference divides by sum_ctx(pc) where pc are RAW (possibly negative) confidences
    (blocks.py:264-267), so |sum pc| << sum |pc| makes that joint ill-conditioned in ANY precision."""
    hs = h[..., :num_joints]
    hc = h[..., num_joints:]
    ps = softargmax_2d_model(ops, hs)
    pc = softargmax_2d_model(ops, hc)
    vc = joints_probability_model(ops, hc)
    if debug is not None:
        v = np.asarray(ops.to_numpy(vc), dtype=np.float64).reshape(-1, num_joints, num_context)
        debug.setdefault('ctx_cond', []).append(np.abs(v).sum(-1) / np.maximum(np.abs(v.sum(-1)), 1e-300))
    pose = context_aggregation_model(ops, ps, pc, vc, num_joints, num_context, alpha)
    visible = joints_probability_model(ops, hs)
    return pose, visible, hs


def pose_regression_2d(ops, h):
    """reception.py:185-190."""
    return softargmax_2d_model(ops, h), joints_probability_model(ops, h), h


def pose_regression_3d(ops, h, num_joints, depth_maps):
    """reception.py:193-222.  Channel c = d*num_joints + j (depth-major)."""
    n, hh, ww, ch = h.shape
    assert ch == depth_maps * num_joints
    h5 = h.reshape(n, hh, ww, depth_maps, num_joints)
    hxy = ops.mean(h5, 3)
    hz = ops.mean(h5, (1, 2))
    pxy = softargmax_2d_model(ops, hxy)
    pz = softargmax_1d_model(ops, hz)
    pose = ops.concat([pxy, pz])
    vxy = ops.amax(hxy, (1, 2))
    vz = ops.amax(hz, 1)
    visible = ops.sigmoid(vxy + vz)[..., None]
    return pose, visible, hxy


def forward(ops, weight_table, x, num_joints, dim, num_context_per_joint=None, alpha=0.8,
            num_blocks=4, depth_maps=16, ksize=(3, 3), export_heatmaps=False,
            concat_pose_confidence=True, return_weights_used=False, debug=None):
    """reception.py:225-319 `build(...)` applied to batch x (N,256,256,3)."""
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

    w = Weights(weight_table, ops)
    c = Ctx(ops, w)
    x = ops.from_numpy(x)
    outputs = []
    x = _stem(c, x)

    for bidx in range(num_blocks):
        nfilt = x.shape[-1]
        x = _reception_block(c, x, 'rBlock%d' % (bidx + 1), ksize)
        ident_map = x
        x = c.sub('SepConv%d' % (bidx + 1)).separable_act_conv_bn(x, nfilt, ksize)
        h = c.sub('RegMap%d' % (bidx + 1)).act_conv(x, num_heatmaps, (1, 1))

        if dim == 2:
            if num_context_per_joint is not None:
                pose, visible, hm = pose_regression_2d_context(
                    ops, h, num_joints, num_context_per_joint, alpha, debug)
            else:
                pose, visible, hm = pose_regression_2d(ops, h)
        else:
            pose, visible, hm = pose_regression_3d(ops, h, num_joints, depth_maps)

        if concat_pose_confidence:
            outputs.append(ops.concat([pose, visible]))
        else:
            outputs.append(pose)
            outputs.append(visible)
        if export_heatmaps:
            outputs.append(hm)

        if bidx < num_blocks - 1:
            h = c.sub('fReMap%d' % (bidx + 1)).act_conv_bn(h, nfilt, (1, 1))
            x = ident_map + x + h

    outputs = [ops.to_numpy(o) for o in outputs]
    if return_weights_used:
        return outputs, w.used
    return outputs
