"""Namespace package: the B200-native JustPIC hot path lives in ``justpic.jl_b200``."""
