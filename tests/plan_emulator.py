"""CPU emulation of a compiled plan -- TEST INFRASTRUCTURE (never imported by the product).

`deephar_b200.compiler.compile_graph` turns a layer graph into a sequence of kernel ops over a planned set of buffers:
BatchNormalization / ReLU folded into conv prologues and epilogues, residual adds (one of them through a 2x upsampling) and
a pooled second output fused into conv kernels, concatenations turned into channel-offset views or copies, soft-max heads
fused into one op, buffers aliased by liveness.  On the GPU that plan is executed by `Model._bind` / `_issue` through the C
ABI.  Here the SAME plan is executed on the CPU, op by op, on numpy float64 buffers laid out exactly as the plan says
(physical slot, per-item stride, channel offset, leading dimension), each op following the contract written in
include/deephar_b200.h with the oracle's primitives (`oracle/ops_np.py`) doing the arithmetic.  Slots start NaN-filled, so
a read of memory no launch has written, or a buffer recycled while still live, poisons the outputs.

What this checks without a GPU: the compiler's fusion decisions, operand order, attribute plumbing, which BatchNormalization
folds into which conv, concat / slice / clip views and the liveness-based slot reuse -- for any model, including the
ones recorded from Keras-style code.  What it does not check: the CUDA kernels (the `-m gpu` tests do).
"""
import numpy as np

from oracle import ops_np as O


def _bn_fold(hw, bn):
    """(scale, shift) of an inference BatchNormalization in float64 (the product folds the same way and rounds the result
    to fp32 for the device, `weights.fold_batchnorm`; tests/test_plan_emulator.py compares the two)."""
    w = bn['weights']
    scale = 1.0 / np.sqrt(hw[w['var']] + O.EPS_BN)
    if 'gamma' in w:
        scale = scale * hw[w['gamma']]
    return scale, hw[w['beta']] - hw[w['mean']] * scale


class PlanEmulator(object):
    def __init__(self, model):
        self.model, self.plan, self.graph = model, model.plan, model.graph
        self.hw = {k: np.asarray(v, np.float64) for k, v in model.get_weights().items()}
        self.launches = 0

    # ---- storage: the plan's own layout -----------------------------------------------------------------------------
    def _items(self, kind):
        return self.n_frames if kind == 'frame' else self.n_frames // self.graph.frames_per_clip

    def _region(self, t):
        s = self.plan.storage[t.id]
        items = self._items(t.kind)
        hw = t.shape[0] * t.shape[1]
        flat = self.slots[s.buf.phys]
        assert items * hw * s.ld <= flat.size, 'tensor %r does not fit its slot' % (t,)
        return flat[:items * hw * s.ld].reshape(items, hw, s.ld), s.c_off

    def get(self, t):
        reg, off = self._region(t)
        a = reg[:, :, off:off + t.shape[2]]
        return a.reshape((a.shape[0],) + tuple(t.shape)).copy()

    def put(self, t, value, c_off=0, channels=None):
        reg, off = self._region(t)
        c = t.shape[2] if channels is None else channels
        value = np.asarray(value, np.float64)
        assert value.size == reg.shape[0] * reg.shape[1] * c, (t, value.shape)
        reg[:, :, off + c_off:off + c_off + c] = value.reshape(reg.shape[0], reg.shape[1], c)

    # ---- ops (include/deephar_b200.h) ---------------------------------------------------------------------------------
    def _conv(self, k, separable):
        a = k.attrs
        x = self.get(k.ins[0])
        if a['pre_bn']:
            sc, sh = _bn_fold(self.hw, a['pre_bn'])
            x = x * sc + sh
        if a['pre_relu']:
            x = np.maximum(x, 0.0)
        if separable:
            y = O.separable_conv2d(x, self.hw[a['depthwise']], self.hw[a['pointwise']], tuple(a['strides']), a['padding'])
        else:
            y = O.conv2d(x, self.hw[a['kernel']], tuple(a['strides']), a['padding'])
        if a['post_bn']:
            sc, sh = _bn_fold(self.hw, a['post_bn'])
            y = y * sc + sh
        if a['post_relu']:
            y = np.maximum(y, 0.0)
        for i in range(a['n_res']):
            r = self.get(k.ins[1 + i])
            if (a.get('res_up2x', 0) >> i) & 1:
                r = O.upsample2d(r)
            y = y + r
        self.put(k.outs[0], y)
        if a.get('pool_out'):
            self.put(k.outs[1], O.maxpool2d(y, (2, 2)))

    def _sam2d(self, k):
        a = k.attrs
        p = O.channel_softmax_2d(self.get(k.ins[0]), a['alpha'])
        pose = O.softargmax2d(p)
        if a['depth']:
            d = self.get(k.ins[1])
            z = np.sum(O.sigmoid(d) * p, axis=(1, 2))[..., None]             # spnet.py:201-205
            pose = np.concatenate([pose, z], axis=-1)
        self.put(k.outs[0], pose)
        self.put(k.outs[1], O.keypoint_confidence(p))
        if a['prob']:
            self.put(k.outs[2], p)

    def _pose3d(self, k, vis_scale=1.0, prob=False):
        a = k.attrs
        h = self.get(k.ins[0])
        n, hh, ww, ch = h.shape
        h5 = h.reshape(n, hh, ww, a['depth_maps'], a['num_joints'])
        hxy, hz = h5.mean(axis=3), h5.mean(axis=(1, 2))
        pose = np.concatenate([O.softargmax2d(O.channel_softmax_2d(hxy)), O.lin_interpolation_1d(O.channel_softmax_1d(hz))],
                              axis=-1)
        vis = O.sigmoid(vis_scale * (hxy.max(axis=(1, 2)) + hz.max(axis=1)))[..., None]
        self.put(k.outs[0], pose)
        self.put(k.outs[1], vis)
        if prob:
            self.put(k.outs[2], O.channel_softmax_2d(hxy))

    def _step(self, k):
        kd, a = k.kind, k.attrs
        if kd == 'conv':
            self._conv(k, False)
        elif kd == 'sepconv':
            self._conv(k, True)
        elif kd == 'maxpool':
            self.put(k.outs[0], O.maxpool2d(self.get(k.ins[0]), tuple(a['pool']), tuple(a['strides']), a['padding']))
        elif kd == 'upsample_add':
            self.put(k.outs[0], self.get(k.ins[0]) + O.upsample2d(self.get(k.ins[1])))
        elif kd == 'upsample':
            self.put(k.outs[0], O.upsample2d(self.get(k.ins[0])))
        elif kd in ('add', 'affine', 'copy'):
            y = sum(self.get(t) for t in k.ins)
            if kd == 'affine':
                if a['bn']:
                    sc, sh = _bn_fold(self.hw, a['bn'])
                    y = y * sc + sh
                if a['relu']:
                    y = np.maximum(y, 0.0)
            if kd == 'copy':
                self.put(k.outs[0], y, c_off=a['c_off'], channels=a['channels'])
            else:
                self.put(k.outs[0], y)
        elif kd == 'scale':
            self.put(k.outs[0], self.get(k.ins[0]) * float(a['value']))
        elif kd == 'pose_regression_2d_context':
            h = self.get(k.ins[0])
            nj, nc = a['num_joints'], a['num_context']
            hs, hc = h[..., :nj], h[..., nj:]
            ys, yc = O.softargmax2d(O.channel_softmax_2d(hs)), O.softargmax2d(O.channel_softmax_2d(hc))
            pc = O.keypoint_confidence(hc)                                   # on RAW maps (blocks.py:328-343)
            grp = lambda v: v.reshape(v.shape[0], nj, nc, -1).sum(axis=2)     # noqa: E731   blocks.py:227-233
            self.put(k.outs[0], a['alpha'] * ys + (1 - a['alpha']) * grp(yc * pc) / grp(pc))
            self.put(k.outs[1], O.keypoint_confidence(hs))
        elif kd == 'pose_regression_2d':
            h = self.get(k.ins[0])
            self.put(k.outs[0], O.softargmax2d(O.channel_softmax_2d(h)))
            self.put(k.outs[1], O.keypoint_confidence(h))
        elif kd == 'pose_regression_3d':
            self._pose3d(k)
        elif kd == 'pose_regression_3d_ex':
            self._pose3d(k, vis_scale=a['vis_scale'], prob=True)
        elif kd == 'sam2d':
            self._sam2d(k)
        elif kd == 'kron':
            p, z = self.get(k.ins[0]), self.get(k.ins[1])
            self.put(k.outs[0], np.einsum('nhwj,nhwf->njf', p, z))
        elif kd == 'mask_mul':
            self.put(k.outs[0], self.get(k.ins[0]) * self.get(k.ins[1]))
        elif kd == 'zeropad':
            self.put(k.outs[0], O.zeropad2d(self.get(k.ins[0]), a['pads']))
        elif kd == 'maxminpool':
            self.put(k.outs[0], O.max_min_pooling(self.get(k.ins[0]), (2, 2), 'same'))
        elif kd == 'global_maxmin_softmax':
            self.put(k.outs[0], O.softmax(O.global_max_min_pooling(self.get(k.ins[0]))))
        else:
            raise NotImplementedError('plan emulator: kernel op %s' % kd)
        self.launches += 1

    def run(self, x):
        """x: (N,H,W,3) frames or (B,T,H,W,3) clips -> the model's outputs in Keras shapes (float64)."""
        x = np.asarray(x, np.float64)
        self.n_frames = int(np.prod(x.shape[:-3]))
        self.slots = [np.full(self._items(kind) * fl, np.nan) for (kind, fl) in self.plan.phys]
        t_in = self.graph.inputs[0]
        self.put(t_in, x.reshape((self.n_frames,) + tuple(t_in.shape)))
        for k in self.plan.kops:
            self._step(k)
        outs = []
        for t in self.graph.outputs:
            items = self._items(t.kind) if t.kind == 'clip' else self.n_frames
            outs.append(self.get(t).reshape(self.model._keras_shape(t, items)))
        return outs
