def unit_norm(*a, **k):
    return None
