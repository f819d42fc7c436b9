"""Statistical screen for candidate counter-based dropout generators (CPU, numpy): the generator must give two 16-bit
lanes per 32-bit output whose keep decisions are unbiased and uncorrelated across lanes, index lags (neighbours, rows,
power-of-two strides) and keys. The shipped generator is `murmur` (vq_hash_pair, csrc/common.cuh: 9 integer instructions
per pair); `mulxor2` (two 32x32->64 multiply + fold rounds) passes the same screen, a single round (`mulxor1`) or a
two-round multiply-xorshift without the third mixing step (`lite`) fail it. Compiled for sm_100a, `mulxor2` saved only 16 of
the 1 872 static SASS instructions of the encoder attention forward (the IMAD.WIDE pairs bring register moves), so the
murmur3 finaliser stays; a real saving needs fewer hashes per element, not a cheaper mix.

    python tools/dropout_hash_eval.py
"""
import numpy as np

THR = round(0.1 * 65536)
M32 = 0xFFFFFFFF


def murmur(key, idx):
    x = (idx * 0x9E3779B1 + key) & M32
    x ^= x >> 16; x = (x * 0x85EBCA6B) & M32
    x ^= x >> 13; x = (x * 0xC2B2AE35) & M32
    x ^= x >> 16
    return x


def lite(key, idx):
    x = (idx * 0x9E3779B1 + key) & M32
    x ^= x >> 16; x = (x * 0x21F0AAAD) & M32
    x ^= x >> 15
    return x


def _fold(x, m):
    p = x * np.uint64(m)          # 32 x 32 -> 64 bits
    return ((p >> 32) ^ (p & M32)) & M32


def mulxor1(key, idx):
    return _fold((idx * 0x9E3779B1 + key) & M32, 0x7F4A7C15)


def mulxor2(key, idx):
    return _fold(_fold((idx ^ key) & M32, 0x4A39B70D), 0x12FAD5C9)


def screen(h, n=1 << 22, n_keys=3):
    idx = np.arange(n, dtype=np.uint64)
    keys = [np.uint64(k) for k in np.random.default_rng(1).integers(0, 2 ** 32, n_keys, dtype=np.uint64)]
    worst, rates, his = 0.0, [], []
    for key in keys:
        x = h(key, idx)
        lo = ((x & 0xFFFF) >= THR).astype(np.float64)
        hi = ((x >> 16) >= THR).astype(np.float64)
        his.append(hi)
        rates += [lo.mean(), hi.mean()]
        worst = max(worst, abs(np.corrcoef(lo, hi)[0, 1]))
        for lag in (1, 2, 3, 4, 8, 16, 32, 64, 384, 768, 1536, 4096, 1 << 15, 1 << 16, 1 << 17, 1 << 20):
            for a, b in ((lo, lo), (hi, hi), (lo, hi)):
                worst = max(worst, abs(np.corrcoef(a[:-lag], b[lag:])[0, 1]))
    cross = abs(np.corrcoef(his[0], his[1])[0, 1])
    return min(rates), max(rates), worst, cross, 3 / np.sqrt(n)


if __name__ == "__main__":
    for name, h in (("murmur", murmur), ("lite", lite), ("mulxor1", mulxor1), ("mulxor2", mulxor2)):
        lo, hi, worst, cross, noise = screen(h)
        ok = worst < 1.5 * noise and cross < 1.5 * noise and abs(lo - 0.9) < 1e-3 and abs(hi - 0.9) < 1e-3
        print(f"{name:8s} keep {lo:.4f}..{hi:.4f}  worst |corr| over lanes/lags {worst:.4f}  across keys {cross:.4f}  "
              f"(3 sigma = {noise:.4f})  {'PASS' if ok else 'FAIL'}")
