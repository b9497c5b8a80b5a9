"""Helpers for hand-scripted playouts on an injected word stream (replay mode)."""
import numpy as np


class Script:
    """Builds the 32-bit words that make java.util.Random's derivations return chosen values."""

    def __init__(self):
        self.words = []

    def int(self, bound: int, value: int):
        assert 0 <= value < bound
        if bound & (bound - 1) == 0:                  # (bound * next(31)) >> 31
            u = value * ((1 << 31) // bound)
        else:                                          # next(31) % bound
            u = value
        self.words.append((u << 1) & 0xFFFFFFFF)
        return self

    def double_k(self, k: int):
        """nextDouble() == k * 2^-53"""
        self.words += [((k >> 27) << 6) & 0xFFFFFFFF, ((k & ((1 << 27) - 1)) << 5) & 0xFFFFFFFF]
        return self

    def double(self, x: float):
        return self.double_k(int(x * (1 << 53)))

    def array(self, pad: int = 0) -> np.ndarray:
        return np.array(self.words + [0] * pad, dtype=np.uint32)


def orx_mix(t: int, owner: int, armies: int) -> int:
    M = 0xFFFFFFFF
    x = ((armies * 2654435761) & M) ^ (((owner + 1) * 0x85EBCA77) & M) ^ (((t + 1) * 0xC2B2AE3D) & M)
    x ^= x >> 15
    x = (x * 0x2C1B3C6D) & M
    x ^= x >> 12
    return x


def board_hash(owner, armies) -> int:
    return sum(orx_mix(t, o & 0xFF, a) for t, (o, a) in enumerate(zip(owner, armies))) & 0xFFFFFFFF
