"""Throughput schedule of one VNLB step (placeholder: delegates to the parity schedule)."""
from .proc_nl import proc_nl


def proc_nl_fast(images, flows, args, stats=None, y_range=None):
    return proc_nl(images, flows, args, stats, y_range)
