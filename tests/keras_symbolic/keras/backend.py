"""keras.backend: the functions the reference's graph-building code and Lambda bodies call on symbolic tensors
(deephar_b200.keras_trace records them); anything else fails when it is called."""
from deephar_b200.keras_trace import (clip, epsilon, exp, expand_dims, image_data_format, int_shape, max, mean,  # noqa: F401,A004
                                      ndim, reshape, set_image_data_format, squeeze, stop_gradient, sum, tile)


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)

    def stub(*args, **kwargs):
        raise NotImplementedError('keras.backend.%s is not part of the recording API (deephar_b200.keras_trace)' % name)
    return stub
