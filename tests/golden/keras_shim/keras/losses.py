def binary_crossentropy(*a, **k):
    raise NotImplementedError


def categorical_crossentropy(*a, **k):
    raise NotImplementedError


def mean_squared_error(*a, **k):
    raise NotImplementedError
