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
  const double kp1 = k + 1.0;
  const double kp1sq = kp1 * kp1;
  return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / kp1;
}

// Bin(n, p) for 0 < p <= 1/2 and n*p >= 10.
B2_HD int64_t binomial_btrs(int64_t n, double p, const Stream& s) {
  const double nd = (double)n;
  const double spq = sqrt(nd * p * (1.0 - p));
  const double b = 1.15 + 2.53 * spq;
  const double a = -0.0873 + 0.0248 * b + 0.01 * p;
  const double c = nd * p + 0.5;
  const double v_r = 0.92 - 4.2 / b;
  const double alpha = (2.83 + 5.1 / b) * spq;
  const double r = p / (1.0 - p);
  const double m = floor((nd + 1.0) * p);
  for (uint32_t attempt = 0;; ++attempt) {
    const Words4 w = stream_block(s, attempt);
    const double u = uniform53(w.w[0], w.w[1]) - 0.5;
    double v = uniform53(w.w[2], w.w[3]);
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + c);
    if (us >= 0.07 && v <= v_r) return (int64_t)k;   // squeeze: ~86 % of attempts
    if (k < 0.0 || k > nd) continue;
    v = log(v * alpha / (a / (us * us) + b));
    const double bound = (m + 0.5) * log((m + 1.0) / (r * (nd - m + 1.0))) +
                         (nd + 1.0) * log((nd - m + 1.0) / (nd - k + 1.0)) +
                         (k + 0.5) * log(r * (nd - k + 1.0) / (k + 1.0)) + stirling_tail(m) +
                         stirling_tail(nd - m) - stirling_tail(k) - stirling_tail(nd - k);
    if (v <= bound) return (int64_t)k;
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

}  // namespace b200boot
