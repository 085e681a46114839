"""Mirror of deephar/models/common.py (forward parts): residual / downscaling / upscaling units
recorded as graph layers (the compiler fuses each unit into one or two kernels)."""
from .layers import (BatchNormalization, add, appstr, concatenate, conv2d, maxpooling2d, relu,
                     sepconv2d, upsampling2d)


def _merge_list(merge, tensors):
    """One tensor passes through; several are merged (common.py:9-22: the reference's two list helpers)."""
    if not isinstance(tensors, list):
        raise AssertionError('t should be a list, got ({})'.format(tensors))
    return merge(tensors) if len(tensors) > 1 else tensors[0]


def concat_tensorlist(t):
    return _merge_list(concatenate, t)


def add_tensorlist(t):
    return _merge_list(add, t)


def residual_unit(x, kernel_size, strides=(1, 1), out_size=None,
                  convtype='depthwise', shortcut_act=True,
                  features_div=2, name=None):
    """(Separable) Residual Unit -- common.py:25-67.  BatchNormalization here is the Keras default
    (scale=True).  Two shapes of the same unit:
      * shape-preserving: shortcut = the raw input, body = BN -> ReLU -> conv;
      * projecting (width or stride changes): BN first, shortcut = 1x1 conv of its (activated) output,
        body = ReLU -> conv on the same normalised tensor.
    convtype 'depthwise': one separable conv; 'normal': 1x1 bottleneck (width / features_div) -> BN -> ReLU -> kxk."""
    assert convtype in ['depthwise', 'normal'], 'Invalid convtype ({}).'.format(convtype)
    width = x.channels
    out_size = width if out_size is None else out_size
    projecting = width != out_size or tuple(strides) != (1, 1)

    normed = BatchNormalization(x, name=appstr(name, '_bn1'))
    if projecting:
        shortcut = relu(normed, name=appstr(name, '_shortcut_act')) if shortcut_act else normed
        shortcut = conv2d(shortcut, out_size, (1, 1), strides=strides, name=appstr(name, '_shortcut_conv'))
    else:
        shortcut = x
    body = relu(normed, name=appstr(name, '_act1'))
    if convtype == 'depthwise':
        body = sepconv2d(body, out_size, kernel_size, strides=strides, name=appstr(name, '_conv1'))
    else:
        body = conv2d(body, int(out_size / features_div), (1, 1), name=appstr(name, '_conv1'))
        body = relu(BatchNormalization(body, name=appstr(name, '_bn2')), name=appstr(name, '_act2'))
        body = conv2d(body, out_size, kernel_size, strides=strides, name=appstr(name, '_conv2'))
    return add([shortcut, body])


def _rescaling_unit(resample, x, cfg, out_size, name):
    if cfg.downsampling_type != 'maxpooling':
        raise NotImplementedError("downsampling_type='conv' is not used by the reference scripts")
    return residual_unit(resample(x, (2, 2)), cfg.kernel_size, out_size=x.channels if out_size is None else out_size,
                         name=appstr(name, '_r0'))


def downscaling_unit(x, cfg, out_size=None, name=None):
    """common.py:70-86 (downsampling_type 'maxpooling'; 'conv' is never used by a shipped script)."""
    return _rescaling_unit(maxpooling2d, x, cfg, out_size, name)


def upscaling_unit(x, cfg, out_size=None, name=None):
    """common.py:89-108."""
    return _rescaling_unit(upsampling2d, x, cfg, out_size, name)


# Aliases (common.py:159-162).
residual = residual_unit
downscaling = downscaling_unit
upscaling = upscaling_unit
