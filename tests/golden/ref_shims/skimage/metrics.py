def peak_signal_noise_ratio(*a, **k):
    raise NotImplementedError("stub")


def structural_similarity(*a, **k):
    raise NotImplementedError("stub")
