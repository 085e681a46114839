"""keras.backend: the few functions the reference's graph-building code calls on symbolic tensors."""


def image_data_format():
    return 'channels_last'


def set_image_data_format(fmt):
    assert fmt == 'channels_last'


def int_shape(x):
    return (None,) + tuple(x.shape)


def ndim(x):
    return len(x.shape) + 1          # + the batch axis (clip inputs are already folded into frames)


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)

    def stub(*args, **kwargs):
        raise NotImplementedError('keras.backend.%s: tensor arithmetic is not part of the recording API' % name)
    return stub
