"""keras.layers -> deephar_b200.keras_compat; every other layer class the reference merely imports is a stub that
fails when it is constructed."""
from deephar_b200.keras_compat import (Activation, Add, AveragePooling2D, BatchNormalization, Concatenate,  # noqa: F401
                                       Conv2D, GlobalMaxPooling1D, GlobalMaxPooling2D, Input, Lambda, MaxPooling2D,
                                       Multiply, SeparableConv2D, TimeDistributed, UpSampling2D, ZeroPadding2D, add,
                                       concatenate, multiply)


def __getattr__(name):
    if name.startswith('__'):
        raise AttributeError(name)

    def stub(*args, **kwargs):
        raise NotImplementedError('keras.layers.%s is not part of the recording API (deephar_b200.keras_compat)' % name)
    stub.__name__ = name
    return stub
