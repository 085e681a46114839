def l1(*a, **k):
    return None


def l2(*a, **k):
    return None
