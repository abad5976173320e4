// Device statistic registry: every moment statistic is (a) a per-row
// contribution vector, (b) a finisher from summed contributions and a count.
// The same two functions give the resample statistic (sums over drawn rows,
// count N), the statistic on the data itself (sums over all rows, count N;
// bootstrap.py:580) and the leave-one-out statistic (totals minus row i,
// count N-1; bootstrap.py:595-599) — so the three are consistent by construction.
#pragma once
#include <math.h>
#include <stdint.h>

#include "philox.cuh"

namespace b200boot {

enum MomentSet : int {
  kMomSum = 0,   // 1 array : {x}
  kMomSum2 = 1,  // 1 array : {x, x^2}            (rows pre-centred)
  kMomAB = 2,    // 2 arrays: {a, b}
  kMomXW = 3,    // 2 arrays: {x*w, w}
  kMomPair5 = 4  // 2 arrays: {a, b, a^2, b^2, a*b} (rows pre-centred)
};

template <int MS> struct MomTraits;
template <> struct MomTraits<kMomSum>   { static constexpr int kArrays = 1, kMoments = 1, kWays = 8; };
template <> struct MomTraits<kMomSum2>  { static constexpr int kArrays = 1, kMoments = 2, kWays = 4; };
template <> struct MomTraits<kMomAB>    { static constexpr int kArrays = 2, kMoments = 2, kWays = 4; };
template <> struct MomTraits<kMomXW>    { static constexpr int kArrays = 2, kMoments = 2, kWays = 4; };
template <> struct MomTraits<kMomPair5> { static constexpr int kArrays = 2, kMoments = 5, kWays = 2; };

constexpr int kMaxMoments = 5;

// contribution of one row; v[a] is the row's value in array a
template <int MS>
B2_HD void contribution(const double* v, double* c) {
  if (MS == kMomSum) {
    c[0] = v[0];
  } else if (MS == kMomSum2) {
    c[0] = v[0];
    c[1] = v[0] * v[0];
  } else if (MS == kMomAB) {
    c[0] = v[0];
    c[1] = v[1];
  } else if (MS == kMomXW) {
    c[0] = v[0] * v[1];
    c[1] = v[1];
  } else {
    c[0] = v[0];
    c[1] = v[1];
    c[2] = v[0] * v[0];
    c[3] = v[1] * v[1];
    c[4] = v[0] * v[1];
  }
}

// stat ids mirror b200boot_stat in include/b200boot.h
enum StatId : int {
  kStatMean = 0, kStatVar = 1, kStatStd = 2, kStatRatio = 3, kStatWavg = 4,
  kStatPearson = 5, kStatMedian = 6, kStatSum = 7,
  kStatSlope = 8,       // least-squares line b ~ slope * a + intercept through paired (a, b):
  kStatIntercept = 9,   //   np.polyfit(a, b, 1) = [slope, intercept]
  kStatLinFit = 10      // both, as a statistic with two outputs
};

inline int moment_set_of(int stat) {
  switch (stat) {
    case kStatMean: case kStatSum: return kMomSum;
    case kStatVar: case kStatStd: return kMomSum2;
    case kStatRatio: return kMomAB;
    case kStatWavg: return kMomXW;
    case kStatPearson: case kStatSlope: case kStatIntercept: case kStatLinFit: return kMomPair5;
    default: return -1;
  }
}
inline bool stat_needs_centering(int stat) {
  return stat == kStatVar || stat == kStatStd || moment_set_of(stat) == kMomPair5;
}
inline int stat_arrays(int stat) {
  return (stat == kStatRatio || stat == kStatWavg || moment_set_of(stat) == kMomPair5) ? 2 : 1;
}
// outputs per resample, and the scalar statistic behind output o
B2_HD int stat_outputs(int stat) { return stat == kStatLinFit ? 2 : 1; }
B2_HD int stat_component(int stat, int o) { return stat == kStatLinFit ? kStatSlope + o : stat; }

// statistic from summed contributions m[] over `cnt` rows; `center` = the per-array means the
// rows were pre-centred on (only the intercept needs them back)
B2_HD double finish_stat(int stat, const double* m, double cnt, const double* center = nullptr) {
  switch (stat) {
    case kStatMean: return m[0] / cnt;
    case kStatSum: return m[0];
    case kStatVar: case kStatStd: {
      // rows are pre-centred on the data mean, so mu is small and the subtraction is
      // benign; what is left below the rounding floor of the two terms (a resample of
      // identical rows) is noise and is returned as the exact 0 numpy's two-pass gives
      const double mu = m[0] / cnt;
      double v = m[1] / cnt - mu * mu;
      if (v <= 8.0 * 2.220446049250313e-16 * mu * mu) v = 0.0;
      if (stat == kStatVar) return v;
      return sqrt(v);
    }
    case kStatRatio: return (m[0] / cnt) / (m[1] / cnt);
    case kStatWavg: return m[0] / m[1];
    case kStatPearson: {
      double saa = m[2] - m[0] * m[0] / cnt;
      double sbb = m[3] - m[1] * m[1] / cnt;
      const double sab = m[4] - m[0] * m[1] / cnt;
      if (saa <= 8.0 * 2.220446049250313e-16 * m[0] * m[0] / cnt) saa = 0.0;   // constant resample: r undefined
      if (sbb <= 8.0 * 2.220446049250313e-16 * m[1] * m[1] / cnt) sbb = 0.0;
      return sab / sqrt(saa * sbb);
    }
    case kStatSlope: case kStatIntercept: {
      double saa = m[2] - m[0] * m[0] / cnt;
      const double sab = m[4] - m[0] * m[1] / cnt;
      if (saa <= 8.0 * 2.220446049250313e-16 * m[0] * m[0] / cnt) saa = 0.0;   // constant abscissa: no line
      const double slope = sab / saa;
      if (stat == kStatSlope) return slope;
      const double ca = center ? center[0] : 0.0, cb = center ? center[1] : 0.0;
      return (cb + m[1] / cnt) - slope * (ca + m[0] / cnt);
    }
    default: return NAN;
  }
}

}  // namespace b200boot
