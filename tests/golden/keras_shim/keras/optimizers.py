class _Opt(object):
    def __init__(self, *a, **k):
        pass


class SGD(_Opt):
    pass


class RMSprop(_Opt):
    pass


class Adam(_Opt):
    pass
