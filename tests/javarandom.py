"""java.util.Random's documented LCG (Java SE API spec), used to pin the oracle's derivations against real Java
outputs and to reproduce the (a, d) pairs that the reference's JUnit tests draw (BattleSimTests.java)."""
import numpy as np

MASK = (1 << 48) - 1


class JavaRandom:
    def __init__(self, seed: int):
        self.seed = (seed ^ 0x5DEECE66D) & MASK

    def next(self, bits: int) -> int:
        self.seed = (self.seed * 0x5DEECE66D + 0xB) & MASK
        v = self.seed >> (48 - bits)
        if bits == 32 and v >= 1 << 31:  # (int) of a 32-bit value; fewer bits are always non-negative
            v -= 1 << 32
        return v

    def next_word(self) -> int:
        """next(32) as an unsigned 32-bit word: the injected-stream word that makes the oracle equal to this LCG"""
        return self.next(32) & 0xFFFFFFFF

    def nextInt(self, bound: int = None) -> int:
        if bound is None:
            return self.next(32)
        r = self.next(31)
        m = bound - 1
        if bound & m == 0:
            return (bound * r) >> 31
        u = r
        while True:
            r = u % bound
            if u - r + m < (1 << 31):
                return r
            u = self.next(31)

    def nextDouble(self) -> float:
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))


def lcg_words(seed: int, n: int) -> np.ndarray:
    r = JavaRandom(seed)
    return np.array([r.next_word() for _ in range(n)], dtype=np.uint32)
