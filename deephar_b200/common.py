"""Mirror of deephar/models/common.py (forward parts): residual / downscaling / upscaling units
recorded as graph layers (the compiler fuses each unit into one or two kernels)."""
from .layers import (BatchNormalization, add, appstr, concatenate, conv2d, maxpooling2d, relu,
                     sepconv2d, upsampling2d)


def concat_tensorlist(t):
    """common.py:9-14."""
    assert isinstance(t, list), 't should be a list, got ({})'.format(t)
    if len(t) > 1:
        return concatenate(t)
    return t[0]


def add_tensorlist(t):
    """common.py:17-22."""
    assert isinstance(t, list), 't should be a list, got ({})'.format(t)
    if len(t) > 1:
        return add(t)
    return t[0]


def residual_unit(x, kernel_size, strides=(1, 1), out_size=None,
                  convtype='depthwise', shortcut_act=True,
                  features_div=2, name=None):
    """(Separable) Residual Unit -- common.py:25-67.  BatchNormalization here is the Keras
    default (scale=True)."""
    assert convtype in ['depthwise', 'normal'], 'Invalid convtype ({}).'.format(convtype)

    num_filters = x.channels
    if out_size is None:
        out_size = num_filters

    skip_conv = (num_filters != out_size) or (tuple(strides) != (1, 1))

    if skip_conv:
        x = BatchNormalization(x, name=appstr(name, '_bn1'))

    shortcut = x
    if skip_conv:
        if shortcut_act:
            shortcut = relu(shortcut, name=appstr(name, '_shortcut_act'))
        shortcut = conv2d(shortcut, out_size, (1, 1), strides=strides,
                          name=appstr(name, '_shortcut_conv'))

    if not skip_conv:
        x = BatchNormalization(x, name=appstr(name, '_bn1'))
    x = relu(x, name=appstr(name, '_act1'))

    if convtype == 'depthwise':
        x = sepconv2d(x, out_size, kernel_size, strides=strides, name=appstr(name, '_conv1'))
    else:
        x = conv2d(x, int(out_size / features_div), (1, 1), name=appstr(name, '_conv1'))
        x = BatchNormalization(x, name=appstr(name, '_bn2'))
        x = relu(x, name=appstr(name, '_act2'))
        x = conv2d(x, out_size, kernel_size, strides=strides, name=appstr(name, '_conv2'))

    x = add([shortcut, x])
    return x


def downscaling_unit(x, cfg, out_size=None, name=None):
    """common.py:70-86 (downsampling_type 'maxpooling'; 'conv' is never used by a shipped script)."""
    if cfg.downsampling_type != 'maxpooling':
        raise NotImplementedError("downsampling_type='conv' is not used by the reference scripts")
    if out_size is None:
        out_size = x.channels
    x = maxpooling2d(x, (2, 2))
    x = residual_unit(x, cfg.kernel_size, out_size=out_size, strides=(1, 1), name=appstr(name, '_r0'))
    return x


def upscaling_unit(x, cfg, out_size=None, name=None):
    """common.py:89-108."""
    if cfg.downsampling_type != 'maxpooling':
        raise NotImplementedError("downsampling_type='conv' is not used by the reference scripts")
    if out_size is None:
        out_size = x.channels
    x = upsampling2d(x, (2, 2))
    x = residual_unit(x, cfg.kernel_size, out_size=out_size, name=appstr(name, '_r0'))
    return x


# Aliases (common.py:159-162).
residual = residual_unit
downscaling = downscaling_unit
upscaling = upscaling_unit
