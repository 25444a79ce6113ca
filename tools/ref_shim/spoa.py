def poa(*a, **k):
    raise RuntimeError("spoa is not available in this container")
