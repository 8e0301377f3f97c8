def estimate_sigma(*a, **k):
    raise NotImplementedError("stub")
