"""Per-layer check of a whole model on the GPU: run the plan kernel by kernel and compare every convolution's
output with the fp64 oracle of THAT layer applied to the inputs the device actually had (prologue BN/ReLU, epilogue
BN/ReLU, residual adds included).  Pinpoints the kernel / shape behind an end-to-end mismatch.

    python tools/layer_check.py reception2d|reception2d_k3|reception3d|spnet_penn|spnet_ntu RES N [repeat]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deephar_b200 import _ffi, reception, spnet  # noqa: E402
from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d  # noqa: E402
from deephar_b200.weights import fold_batchnorm  # noqa: E402
from oracle import ops_np  # noqa: E402


def build(name, res, frames=2):
    if name == 'reception2d':
        return reception.build((res, res, 3), num_joints=16, dim=2, num_context_per_joint=2, num_blocks=2, ksize=(5, 5),
                               concat_pose_confidence=False), False
    if name == 'reception2d_k3':
        return reception.build((res, res, 3), num_joints=16, dim=2, num_blocks=2, ksize=(3, 3), export_heatmaps=True), False
    if name == 'reception3d':
        return reception.build((res, res, 3), num_joints=17, dim=3, num_blocks=2, ksize=(5, 5), concat_pose_confidence=False), False
    if name == 'spnet_penn':
        return spnet.build(ModelConfig((frames, res, res, 3), pa16j2d, num_actions=[15], num_pyramids=2, action_pyramids=[1, 2],
                                       num_levels=4, pose_replica=True, num_pose_features=160, num_visual_features=160)), True
    return spnet.build(ModelConfig((frames, res, res, 3), pa17j3d, num_actions=[60], num_pyramids=2, action_pyramids=[1, 2],
                                   num_levels=4, num_pose_features=192, num_visual_features=192)), True


def main():
    name, res, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    repeat = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    m, clip = build(name, res)
    m.init_synthetic_weights(1234)
    m.use_cuda_graph = False
    T = m.graph.frames_per_clip
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, ((n, T, res, res, 3) if clip else (n, res, res, 3))).astype(np.float32)
    n_frames = n * T if clip else n
    b = m._bind(n_frames)
    hw = m.get_weights()
    plan = m.plan
    s_in = plan.storage[m.graph.inputs[0].id]
    stream = torch.cuda.current_stream().cuda_stream
    m._ctx.set_workspace(b.workspace.data_ptr(), b.workspace.numel() * 4)
    lib = _ffi.lib()

    def fetch(t):
        s = plan.storage[t.id]
        items = m._items(t.kind, n_frames)
        base = b.slots[s.buf.phys].view(items, t.shape[0], t.shape[1], s.ld)
        return base[..., s.c_off:s.c_off + t.shape[2]].double().cpu().numpy()

    def fold(bn):
        w = bn['weights']
        sc, sh = fold_batchnorm(hw[w['gamma']] if 'gamma' in w else None, hw[w['beta']], hw[w['mean']], hw[w['var']])
        return sc.astype(np.float64), sh.astype(np.float64)

    worst = []
    for rep in range(repeat):
        b.slots[s_in.buf.phys].copy_(torch.from_numpy(x).reshape(-1).cuda())
        bad = 0
        for k, call in zip(plan.kops, b.calls):
            ins = [fetch(t) for t in k.ins] if k.kind in ('conv', 'sepconv') else None
            rc = call[1](*call[2:], stream)
            _ffi.check(rc, call[0])
            torch.cuda.synchronize()
            if k.kind not in ('conv', 'sepconv'):
                continue
            path = lib.dh_last_conv_path(m._ctx.handle)
            a = k.attrs
            xin = ins[0]
            if a['pre_bn']:
                sc, sh = fold(a['pre_bn'])
                xin = xin * sc + sh
            if a['pre_relu']:
                xin = np.maximum(xin, 0)
            if k.kind == 'conv':
                ref = ops_np.conv2d(xin, hw[a['kernel']].astype(np.float64), a['strides'], a['padding'])
            else:
                ref = ops_np.separable_conv2d(xin, hw[a['depthwise']].astype(np.float64), hw[a['pointwise']].astype(np.float64),
                                              a['strides'], a['padding'])
            if a['post_bn']:
                sc, sh = fold(a['post_bn'])
                ref = ref * sc + sh
            if a['post_relu']:
                ref = np.maximum(ref, 0)
            for r in ins[1:]:
                ref = ref + r
            got = fetch(k.outs[0])
            err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
            label = '%s %s->%s k%s s%s path=%d' % (k.kind, 'x'.join(map(str, k.ins[0].shape)), 'x'.join(map(str, k.outs[0].shape)),
                                                  'x'.join(map(str, a['size'])), a['strides'][0], path)
            if err > 1e-4 or not np.isfinite(err):
                bad += 1
                where = np.unravel_index(np.argmax(np.abs(got - ref)), got.shape)
                print('rep %d  BAD %-70s err %.3e at %s' % (rep, label, err, where))
            worst.append((err, label))
        print('rep %d: %d convolutions checked, %d bad' % (rep, len([k for k in plan.kops if k.kind in ('conv', 'sepconv')]), bad))
    worst.sort(reverse=True)
    print('largest errors:')
    for e, l in worst[:6]:
        print('   %.3e  %s' % (e, l))


if __name__ == '__main__':
    main()
