// Counter-based binomial sampler used by the tile-occupancy (multinomial) chain.
//
// A bootstrap resample of N rows is N iid uniform row draws
// (/root/reference/src/scikits/bootstrap/bootstrap.py:680).  Cutting the rows
// into cells, the number of draws landing in each cell is multinomial; it is
// sampled as a chain of conditional binomials, after which the draws inside a
// cell are iid uniform on that cell.  The joint law of the resulting multiset
// is exactly that of N iid draws.
//
// Sampler: Hörmann's transformed-rejection BTRS ("The generation of binomial
// random variates", J. Stat. Comput. Simul. 46, 1993) for n*p >= 10, sequential
// inversion below that.  Every uniform comes from Philox keyed by
// (seed, resample, cell, attempt), so the result is a pure function of those.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "philox.cuh"

namespace b200boot {

// log(k!) - [log(sqrt(2*pi)) + (k + 1/2) log(k+1) - (k+1)]
B2_HD double stirling_tail(double k) {
  if (k <= 9.0) {
    switch ((int)k) {
      case 0: return 0.08106146679532733;
      case 1: return 0.04134069595540946;
      case 2: return 0.027677925684997717;
      case 3: return 0.020790672103765395;
      case 4: return 0.016644691189820815;
      case 5: return 0.013876128823072875;
      case 6: return 0.011896709945893313;
      case 7: return 0.010411265261971892;
      case 8: return 0.009255462182707674;
      default: return 0.008330563433357696;
    }
  }
  const double r = 1.0 / (k + 1.0);      // one division; Horner in r^2
  const double r2 = r * r;
  return r * (1.0 / 12 - r2 * (1.0 / 360 - r2 * (1.0 / 1260)));
}

// ---- BTRS in pieces, so the sequential sampler below and the lane state machine of
// the class-split kernel (k1_resample.cuh) run the very same arithmetic ---------------
struct BtrsConsts {
  double nd, b, a, c, v_r, alpha, r, m;
};

// constants of Bin(n, p), 0 < p <= 1/2, n*p >= 10;  r = p / (1 - p) is passed in
B2_HD BtrsConsts btrs_setup(int64_t n, double p, double r) {
  BtrsConsts k;
  k.nd = (double)n;
  const double spq = sqrt(k.nd * p * (1.0 - p));
  k.b = 1.15 + 2.53 * spq;
  k.a = -0.0873 + 0.0248 * k.b + 0.01 * p;
  k.c = k.nd * p + 0.5;
  const double inv_b = 1.0 / k.b;
  k.v_r = 0.92 - 4.2 * inv_b;
  k.alpha = (2.83 + 5.1 * inv_b) * spq;
  k.r = r;
  k.m = floor((k.nd + 1.0) * p);
  return k;
}

// One proposal from the words of attempt block w.  Returns +1 accepted by the squeeze
// (~86 % of proposals), 0 needs the full test (out: k, inv_us, v), -1 rejected (k out
// of range).
B2_HD int btrs_propose(const BtrsConsts& q, const Words4& w, double* k_out, double* inv_us_out,
                       double* v_out) {
  const double u = uniform53(w.w[0], w.w[1]) - 0.5;
  const double v = uniform53(w.w[2], w.w[3]);
  const double us = 0.5 - fabs(u);
  const double inv_us = 1.0 / us;
  const double k = floor((2.0 * q.a * inv_us + q.b) * u + q.c);
  *k_out = k;
  *inv_us_out = inv_us;
  *v_out = v;
  if (us >= 0.07 && v <= q.v_r) return 1;
  if (k < 0.0 || k > q.nd) return -1;
  return 0;
}

// The full acceptance test is  log(v*alpha/(a/us^2+b)) <= head + t2 + t3  with four
// terms of one shape, A*log(B/C) + sg*(tail(D)+tail(E)):
//   term 0 (lhs)  A=1      B=v*alpha      C=a/us^2+b       sg=0
//   term 1 (head) A=m+1/2  B=m+1          C=r(n-m+1)       sg=+1  D=m  E=n-m
//   term 2        A=n+1    B=n-m+1        C=n-k+1          sg=0
//   term 3        A=k+1/2  B=r(n-k+1)     C=k+1            sg=-1  D=k  E=n-k
// Kept as one function so a warp can evaluate the terms of several pending proposals
// side by side (k0_class_counts_kernel) with the arithmetic of the sequential sampler.
B2_HD double btrs_term(int which, double nd, double m, double r, double k, double lhs_num,
                       double lhs_den) {
  double A, B, C, D, E, sg;
  if (which == 0) {
    A = 1.0; B = lhs_num; C = lhs_den; D = 100.0; E = 100.0; sg = 0.0;
  } else if (which == 1) {
    A = m + 0.5; B = m + 1.0; C = r * (nd - m + 1.0); D = m; E = nd - m; sg = 1.0;
  } else if (which == 2) {
    A = nd + 1.0; B = nd - m + 1.0; C = nd - k + 1.0; D = 100.0; E = 100.0; sg = 0.0;
  } else {
    A = k + 0.5; B = r * (nd - k + 1.0); C = k + 1.0; D = k; E = nd - k; sg = -1.0;
  }
  return A * log(B / C) + sg * (stirling_tail(D) + stirling_tail(E));
}

B2_HD void btrs_lhs_args(const BtrsConsts& q, double inv_us, double v, double* num, double* den) {
  *num = v * q.alpha;
  *den = q.a * inv_us * inv_us + q.b;
}

// full acceptance test of proposal k
B2_HD bool btrs_accept(const BtrsConsts& q, double k, double inv_us, double v) {
  double num, den;
  btrs_lhs_args(q, inv_us, v, &num, &den);
  const double t0 = btrs_term(0, q.nd, q.m, q.r, k, num, den);
  const double t1 = btrs_term(1, q.nd, q.m, q.r, k, num, den);
  const double t2 = btrs_term(2, q.nd, q.m, q.r, k, num, den);
  const double t3 = btrs_term(3, q.nd, q.m, q.r, k, num, den);
  return t0 <= (t1 + t2) + t3;
}

// Bin(n, p) for 0 < p <= 1/2 and n*p >= 10.
B2_HD int64_t binomial_btrs(int64_t n, double p, const Stream& s) {
  const BtrsConsts q = btrs_setup(n, p, p / (1.0 - p));
  for (uint32_t attempt = 0;; ++attempt) {
    const Words4 w = stream_block(s, attempt);
    double k, inv_us, v;
    const int verdict = btrs_propose(q, w, &k, &inv_us, &v);
    if (verdict > 0) return (int64_t)k;
    if (verdict == 0 && btrs_accept(q, k, inv_us, v)) return (int64_t)k;
  }
}

// Bin(n, p) for 0 < p <= 1/2 and n*p < 10: walk the cdf from 0.
B2_HD int64_t binomial_inversion(int64_t n, double p, const Stream& s) {
  const double q = 1.0 - p;
  const double ratio = p / q;
  const double g = (double)(n + 1) * ratio;
  const double f0 = exp((double)n * log1p(-p));
  for (uint32_t attempt = 0;; ++attempt) {
    const Words4 w = stream_block(s, attempt);
    double u = uniform53(w.w[0], w.w[1]);
    double f = f0;
    int64_t x = 0;
    bool ok = true;
    while (u > f) {
      u -= f;
      ++x;
      if (x > n || x > 200) {   // rounding pushed u past the tail: redraw
        ok = false;
        break;
      }
      f *= (g / (double)x - ratio);
    }
    if (ok) return x;
  }
}

// Bin(n, num/den) with the probability given as an exact integer ratio.
B2_HD int64_t binomial_ratio(int64_t n, int64_t num, int64_t den, const Stream& s) {
  if (n <= 0 || num <= 0) return 0;
  if (num >= den) return n;
  const bool flip = 2 * num > den;
  const double p = flip ? (double)(den - num) / (double)den : (double)num / (double)den;
  const int64_t k = ((double)n * p >= 10.0) ? binomial_btrs(n, p, s) : binomial_inversion(n, p, s);
  return flip ? n - k : k;
}

// ---- Bin(n, 1/2): the node sampler of the bank-class split tree ---------------------------
// Every split of the class tree (draw.cuh, class_count_tree) is an even split, so one
// sampler with p = 1/2 serves all of them:
//   n <  kHalfBtrsMin: the number of set bits among n random bits (exact, no rejection);
//   n >= kHalfBtrsMin: BTRS specialised to p = 1/2.  Its constants depend on n alone (kernels
//        read them from a table built by half_setup itself, so table and direct evaluation
//        agree bit for bit); the uniforms are built from the mantissa bits directly; and the
//        full acceptance test compares with the exact ratio of the two probabilities,
//        f(k) / f(m) = m! (n-m)! / (k! (n-k)!), through log-factorials (a table of
//        lgamma(j + 1) in the kernels, lgamma itself elsewhere: the same values).
// Explicit fma() where a product feeds a sum, so every caller rounds the same way.
constexpr int64_t kHalfBtrsMin = 20;      // n * p >= 10

struct HalfConsts {
  double b, a, v_r, alpha;
};

B2_HD HalfConsts half_setup(int64_t n) {
  const BtrsConsts q = btrs_setup(n, 0.5, 1.0);
  HalfConsts h;
  h.b = q.b;
  h.a = q.a;
  h.v_r = q.v_r;
  h.alpha = q.alpha;
  return h;
}

B2_HD int popcount32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// [1, 2) from 52 random bits: the top 20 of `hi` and all of `lo` as the mantissa
B2_HD double unit12(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)(0x3FF00000u | (hi >> 12)), (int)lo);
#else
  const uint64_t bits = ((uint64_t)(0x3FF00000u | (hi >> 12)) << 32) | (uint64_t)lo;
  double d;
  memcpy(&d, &bits, sizeof d);
  return d;
#endif
}

// log(j!) for an integer-valued j >= 0
B2_HD double log_factorial(double j) { return lgamma(j + 1.0); }

// n < kHalfBtrsMin
B2_HD int64_t binomial_half_small(int64_t n, const Words4& w) {
  return n <= 0 ? 0 : (int64_t)popcount32(w.w[0] & ((1u << (int)n) - 1u));
}

// One proposal of Bin(n, 1/2), n >= kHalfBtrsMin, from the words of an attempt block:
// +1 accepted by the squeeze, 0 needs half_accept (out: k, inv_us, v), -1 rejected.
B2_HD int half_propose(double nd, const HalfConsts& h, const Words4& w, double* k_out, double* inv_us_out,
                       double* v_out) {
  const double u = unit12(w.w[0], w.w[1]) - 1.5;                          // [-1/2, 1/2)
  const double v = unit12(w.w[2], w.w[3]) - (1.0 - 1.1102230246251565e-16);   // (0, 1): - 1 + 2^-53, exact
  const double us = 0.5 - fabs(u);
  *v_out = v;
  if (us <= 0.0) return -1;                                               // u == -1/2 (probability 2^-52)
  const double inv_us = 1.0 / us;
  const double k = floor(fma(fma(2.0 * h.a, inv_us, h.b), u, fma(nd, 0.5, 0.5)));
  *k_out = k;
  *inv_us_out = inv_us;
  if (us >= 0.07 && v <= h.v_r) return 1;
  if (k < 0.0 || k > nd) return -1;
  return 0;
}

// Full acceptance test of proposal k; lf(j) = log(j!)
template <typename LF>
B2_HD bool half_accept(double nd, const HalfConsts& h, double k, double inv_us, double v, LF&& lf) {
  const double m = floor((nd + 1.0) * 0.5);
  const double lhs = log(v * h.alpha / fma(h.a * inv_us, inv_us, h.b));
  const double rhs = (lf(m) - lf(k)) + (lf(nd - m) - lf(nd - k));
  return lhs <= rhs;
}

// Bin(n, 1/2), n >= kHalfBtrsMin, from attempt `first_attempt` on
template <typename LF>
B2_HD int64_t binomial_half_from(int64_t n, const HalfConsts& h, const Stream& s, uint32_t first_attempt, LF&& lf) {
  const double nd = (double)n;
  for (uint32_t attempt = first_attempt;; ++attempt) {
    const Words4 w = stream_block(s, attempt);
    double k, inv_us, v;
    const int verdict = half_propose(nd, h, w, &k, &inv_us, &v);
    if (verdict > 0) return (int64_t)k;
    if (verdict == 0 && half_accept(nd, h, k, inv_us, v, lf)) return (int64_t)k;
  }
}

B2_HD int64_t binomial_half(int64_t n, const Stream& s) {
  if (n < kHalfBtrsMin) return binomial_half_small(n, stream_block(s, 0u));
  return binomial_half_from(n, half_setup(n), s, 0u, [](double j) { return log_factorial(j); });
}

}  // namespace b200boot
