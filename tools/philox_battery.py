"""Statistical battery for the Philox4x32-R streams as THIS library keys and consumes them
(csrc/philox.cuh counter layout, csrc/draw.cuh 10-bit draws).  It is not TestU01: BigCrush /
PractRand are not available offline.  What it adds to the published result (Salmon et al.,
SC'11: Philox4x32 is Crush-resistant from 7 rounds on) is (a) the same tests run on the
library's own counter patterns — block index, resample id, cell id and seed each swept as
the fast-moving input — and (b) a demonstration that the tests have power: the same battery
rejects Philox4x32 with 2-4 rounds and passes 7 and 10.

    python tools/philox_battery.py [log2_blocks=22] > profiles/r02_philox_battery.log

Every test reports a p-value (two-sided where it makes sense).  A battery of ~60 p-values
should show none below 1e-4; the summary line counts them.
"""
import math
import sys
import time

import numpy as np
from scipy import stats as sps

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
U32 = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox(c0, c1, c2, c3, k0, k1, rounds):
    """Vectorised Philox4x32-`rounds`; inputs broadcastable uint64 arrays holding 32-bit values."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint64)
    k1 = np.asarray(k1, dtype=np.uint64)
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = (p1 >> S32) ^ c1 ^ k0, p1 & U32, (p0 >> S32) ^ c3 ^ k1, p0 & U32
        k0 = (k0 + np.uint64(W0)) & U32
        k1 = (k1 + np.uint64(W1)) & U32
    shape = np.broadcast(c0, c1, c2, c3).shape
    out = np.empty(shape + (4,), dtype=np.uint32)
    for i, c in enumerate((c0, c1, c2, c3)):
        out[..., i] = np.broadcast_to(c, shape).astype(np.uint32)
    return out


def stream(kind, n, rounds, seed=0x1234567):
    """n blocks with ONE input swept 0..n-1 and the others fixed, in the library's layout:
    key = seed, c0 = block, c1 = cell id, c2 = resample id, c3 = purpose << 16."""
    i = np.arange(n, dtype=np.uint64)
    cell = np.uint64(0x80000000 | (37 << 4) | 5)        # class 5 of full tile 37
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)
    z = np.uint64(0)
    if kind == "block":
        return philox(i, cell, np.uint64(4242), z, k0, k1, rounds)
    if kind == "resample":
        return philox(z, cell, i, z, k0, k1, rounds)
    if kind == "cell":
        return philox(z, np.uint64(0x80000000) | i, np.uint64(4242), z, k0, k1, rounds)
    if kind == "seed":
        return philox(z, cell, np.uint64(4242), z, i & U32, i >> S32, rounds)
    raise ValueError(kind)


def chi2_p(obs, exp):
    obs = np.asarray(obs, dtype=float)
    exp = np.broadcast_to(np.asarray(exp, dtype=float), obs.shape)
    x = ((obs - exp) ** 2 / exp).sum()
    dof = obs.size - 1
    return float(sps.chi2.sf(x, dof)) if dof < 5000 else float(sps.norm.sf((x - dof) / math.sqrt(2 * dof)))


def two_sided(p_upper):
    return float(min(1.0, 2 * min(p_upper, 1 - p_upper)))


# ---- tests ---------------------------------------------------------------------------

def t_monobit(w):
    """every one of the 128 output bit positions is set half the time"""
    n = w.shape[0]
    ones = np.array([[(w[:, j] >> np.uint32(b) & np.uint32(1)).sum(dtype=np.int64) for b in range(32)] for j in range(4)])
    z = (ones - n / 2) / math.sqrt(n / 4)
    return float(sps.chi2.sf((z ** 2).sum(), 128))


def t_bytes(w):
    """16 byte positions, 256 cells each; worst p-value, Bonferroni-corrected"""
    n = w.shape[0]
    b = w.view(np.uint8).reshape(n, 16)
    ps = [chi2_p(np.bincount(b[:, j], minlength=256), n / 256) for j in range(16)]
    return float(min(1.0, 16 * min(ps)))


def draws_of(w):
    """the twelve 10-bit row numbers of every block, in the order the kernels consume them"""
    m = np.uint32(1023)
    f = np.stack([(w >> np.uint32(7)) & m, (w >> np.uint32(17)) & m,
                  ((w << np.uint32(5)) | (w >> np.uint32(27))) & m], axis=-1)      # [n][4][3]
    return f.reshape(w.shape[0], 12)


def t_serial_draws(w):
    """pairs of consecutive draws inside a block: 2^20 cells"""
    d = draws_of(w).astype(np.int64)
    pairs = (d[:, 0::2] << 10 | d[:, 1::2]).reshape(-1)
    return chi2_p(np.bincount(pairs, minlength=1 << 20), pairs.size / (1 << 20))


def t_serial_across_blocks(w):
    """last draw of block i with the first draw of block i + 1 (consecutive sweep values)"""
    d = draws_of(w).astype(np.int64)
    pairs = d[:-1, 11] << 10 | d[1:, 0]
    return chi2_p(np.bincount(pairs, minlength=1 << 20), pairs.size / (1 << 20))


def t_draw_uniform(w):
    """each of the 12 draw slots uniform over the 1024 rows; worst, Bonferroni"""
    d = draws_of(w)
    ps = [chi2_p(np.bincount(d[:, j], minlength=1024), d.shape[0] / 1024) for j in range(12)]
    return float(min(1.0, 12 * min(ps)))


def t_birthday(w, m=4096, reps=2048):
    """Marsaglia's birthday spacings on 32-bit words: duplicate spacings ~ Poisson(m^3 / 2^34)"""
    reps = min(reps, w.size // m)
    x = w.reshape(-1)[: m * reps].astype(np.int64).reshape(reps, m)
    x.sort(axis=1)
    sp = np.diff(x, axis=1)
    sp.sort(axis=1)
    dup = (np.diff(sp, axis=1) == 0).sum(axis=1)
    lam = m ** 3 / 2.0 ** 34
    top = 12
    obs = np.bincount(np.minimum(dup, top), minlength=top + 1)
    pk = sps.poisson.pmf(np.arange(top), lam)
    exp = np.append(pk, 1 - pk.sum()) * reps
    return chi2_p(obs, exp)


def t_rank32(w, n_mat=20000):
    """rank over GF(2) of 32 x 32 matrices built from 32 consecutive words"""
    n_mat = min(n_mat, w.size // 32)
    rows = w.reshape(-1)[: 32 * n_mat].astype(np.uint64).reshape(n_mat, 32).copy()
    rank = np.zeros(n_mat, dtype=np.int64)
    for bit in range(32):
        mask = np.uint64(1 << bit)
        has = (rows & mask) != 0                           # [n_mat][32]
        piv = has.argmax(axis=1)
        any_ = has.any(axis=1)
        prow = rows[np.arange(n_mat), piv]
        rank += any_
        elim = has & any_[:, None]
        rows = np.where(elim, rows ^ prow[:, None], rows)
        # the pivot row itself became zero (xor with itself): it is out of the game
    obs = np.array([(rank <= 29).sum(), (rank == 30).sum(), (rank == 31).sum(), (rank == 32).sum()])
    p = np.array([0.0052854502, 0.1283502644, 0.5775761902, 0.2887880951])
    return chi2_p(obs, p * n_mat)


def t_hamming_consecutive(w):
    """Hamming distance between the 128-bit outputs of consecutive sweep values ~ Bin(128, 1/2)"""
    x = (w[1:] ^ w[:-1]).view(np.uint8)
    hd = np.unpackbits(x, axis=1).sum(axis=1, dtype=np.int64) if x.shape[0] <= (1 << 21) else \
        np.unpackbits(x[: 1 << 21], axis=1).sum(axis=1, dtype=np.int64)
    lo, hi = 44, 84
    obs = np.bincount(np.clip(hd, lo, hi) - lo, minlength=hi - lo + 1)
    pk = sps.binom.pmf(np.arange(lo, hi + 1), 128, 0.5)
    pk[0] = sps.binom.cdf(lo, 128, 0.5)
    pk[-1] = sps.binom.sf(hi - 1, 128, 0.5)
    return chi2_p(obs, pk * hd.size)


def t_dft(w, n_bits=1 << 20):
    """NIST spectral test on a bit stream (first output word of every block, bit 13... all bits)"""
    n_bits = min(n_bits, 32 * w.size)
    bits = np.unpackbits(w.reshape(-1)[: n_bits // 32].view(np.uint8)).astype(np.float64) * 2 - 1
    f = np.abs(np.fft.rfft(bits))[: n_bits // 2]
    t = math.sqrt(math.log(1 / 0.05) * n_bits)
    n0 = 0.95 * n_bits / 2
    n1 = (f < t).sum()
    d = (n1 - n0) / math.sqrt(n_bits * 0.95 * 0.05 / 4)
    return float(math.erfc(abs(d) / math.sqrt(2)))


def t_lag_correlation(w):
    """lag-1 and lag-4 correlation of the words as uniforms (word stream order)"""
    u = w.reshape(-1).astype(np.float64) / 2.0 ** 32 - 0.5
    n = u.size
    ps = []
    for lag in (1, 4):
        r = (u[:-lag] * u[lag:]).mean() * 12
        ps.append(two_sided(float(sps.norm.sf(r * math.sqrt(n - lag)))))
    return float(min(1.0, 2 * min(ps)))


def t_avalanche(rounds, n=1 << 15, seed=99):
    """strict avalanche: flipping any one of the 192 input bits flips every output bit with
    probability 1/2; chi-square over the 192 x 128 flip frequencies, and the worst |z|"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 1 << 32, size=(6, n), dtype=np.uint64)
    ref = philox(*base, rounds)
    z2, worst = 0.0, 0.0
    for inp in range(6):
        for bit in range(32):
            mod = base.copy()
            mod[inp] ^= np.uint64(1 << bit)
            d = (philox(*mod, rounds) ^ ref).view(np.uint8).reshape(n, 16)
            flips = np.unpackbits(d, axis=1).sum(axis=0, dtype=np.int64)      # [128]
            z = (flips - n / 2) / math.sqrt(n / 4)
            z2 += (z ** 2).sum()
            worst = max(worst, float(np.abs(z).max()))
    dof = 192 * 128
    return float(sps.norm.sf((z2 - dof) / math.sqrt(2 * dof))), worst


def t_adjacent_streams(rounds, n):
    """two streams that differ by one in the resample id (the library's neighbouring resamples):
    correlation of their uniforms and bit-agreement rate of their words"""
    a = stream("block", n, rounds)
    i = np.arange(n, dtype=np.uint64)
    b = philox(i, np.uint64(0x80000000 | (37 << 4) | 5), np.uint64(4243), np.uint64(0),
               np.uint64(0x1234567), np.uint64(0), rounds)
    ua = a.reshape(-1).astype(np.float64) / 2.0 ** 32 - 0.5
    ub = b.reshape(-1).astype(np.float64) / 2.0 ** 32 - 0.5
    r = (ua * ub).mean() * 12 * math.sqrt(ua.size)
    p_corr = two_sided(float(sps.norm.sf(r)))
    same = 32 * ua.size - int(np.unpackbits((a ^ b).view(np.uint8)).sum(dtype=np.int64))
    zb = (same - 16 * ua.size) / math.sqrt(8 * ua.size)
    return float(min(1.0, 2 * min(p_corr, two_sided(float(sps.norm.sf(zb))))))


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    n = 1 << log2n
    print(f"# Philox4x32-R battery on the library's counter layout; {n} blocks ({4 * n} words) per stream")
    print("# p-values; '<1e-4' marks a failure.  rounds 2-4 show the tests have power; 7 is the library's, 10 Random123's default")
    tests = [("monobit", t_monobit), ("bytes", t_bytes), ("draw_uniform", t_draw_uniform),
             ("serial_draws", t_serial_draws), ("serial_across_blocks", t_serial_across_blocks),
             ("birthday", t_birthday), ("rank32", t_rank32), ("hamming_consecutive", t_hamming_consecutive),
             ("dft", t_dft), ("lag_correlation", t_lag_correlation)]
    kinds = ["block", "resample", "cell", "seed"]
    for rounds in (2, 3, 4, 5, 7, 10):
        t0 = time.time()
        fails, total, lines = 0, 0, []
        nn = n if rounds >= 5 else min(n, 1 << 20)           # the weak ones fail on far less data
        for kind in kinds:
            w = stream(kind, nn, rounds)
            for name, fn in tests:
                p = fn(w)
                total += 1
                fails += p < 1e-4
                lines.append(f"rounds={rounds:2d} stream={kind:9s} {name:22s} p={p:.4g}{'   <1e-4' if p < 1e-4 else ''}")
        p, worst = t_avalanche(rounds)
        total += 1
        fails += p < 1e-4
        lines.append(f"rounds={rounds:2d} inputs=random    avalanche(192x128)     p={p:.4g} worst|z|={worst:.2f}{'   <1e-4' if p < 1e-4 else ''}")
        p = t_adjacent_streams(rounds, min(nn, 1 << 20))
        total += 1
        fails += p < 1e-4
        lines.append(f"rounds={rounds:2d} streams=r,r+1    adjacent_streams       p={p:.4g}{'   <1e-4' if p < 1e-4 else ''}")
        print("\n".join(lines))
        print(f"## rounds={rounds}: {fails} of {total} p-values below 1e-4   ({time.time() - t0:.0f} s)\n", flush=True)


if __name__ == "__main__":
    main()
