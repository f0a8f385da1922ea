// ORACLE — test infrastructure only. Nothing under pytheiasfm_b200/ may include, link or
// call this. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, as the checker and the CPU baseline.
//
// Forward-mode dual numbers. The reference differentiates ReprojectionError with
// ceres::AutoDiffCostFunction (sfm/camera/create_reprojection_error_cost_function.h:63-128),
// i.e. ceres::Jet<double, N>. Ceres is NOT vendored in /root/reference, so the
// Jet algebra is restated here from its published definition: a + v·eps with eps^2 = 0;
// comparisons act on the scalar part only.
#ifndef ORACLE_JET_H_
#define ORACLE_JET_H_

#include <cmath>

namespace oracle {

template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // ceres/jet.h: h = f/g, dh = (df - h dg)/g
  Jet<N> h; const double gi = 1.0 / g.a; h.a = f.a * gi;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi; return h; }

#define ORACLE_JET_MIXED(op)                                                              \
  template <int N> inline Jet<N> operator op(const Jet<N>& f, double s) { return f op Jet<N>(s); } \
  template <int N> inline Jet<N> operator op(double s, const Jet<N>& f) { return Jet<N>(s) op f; }
ORACLE_JET_MIXED(+) ORACLE_JET_MIXED(-) ORACLE_JET_MIXED(*) ORACLE_JET_MIXED(/)
#undef ORACLE_JET_MIXED

template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator-=(Jet<N>& f, const Jet<N>& g) { f = f - g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, const Jet<N>& g) { f = f * g; return f; }
template <int N> inline Jet<N>& operator/=(Jet<N>& f, const Jet<N>& g) { f = f / g; return f; }

#define ORACLE_JET_CMP(op)                                                                \
  template <int N> inline bool operator op(const Jet<N>& f, const Jet<N>& g) { return f.a op g.a; } \
  template <int N> inline bool operator op(const Jet<N>& f, double g) { return f.a op g; }          \
  template <int N> inline bool operator op(double f, const Jet<N>& g) { return f op g.a; }
ORACLE_JET_CMP(<) ORACLE_JET_CMP(<=) ORACLE_JET_CMP(>) ORACLE_JET_CMP(>=) ORACLE_JET_CMP(==) ORACLE_JET_CMP(!=)
#undef ORACLE_JET_CMP

template <int N> inline Jet<N> chain(const Jet<N>& f, double val, double dval) {
  Jet<N> h; h.a = val; for (int i = 0; i < N; ++i) h.v[i] = dval * f.v[i]; return h; }

inline double sqrt_(double x) { return std::sqrt(x); }
inline double sin_(double x) { return std::sin(x); }
inline double cos_(double x) { return std::cos(x); }
inline double tan_(double x) { return std::tan(x); }
inline double atan_(double x) { return std::atan(x); }
inline double abs_(double x) { return std::fabs(x); }
inline double atan2_(double y, double x) { return std::atan2(y, x); }
inline double scalar(double x) { return x; }

template <int N> inline Jet<N> sqrt_(const Jet<N>& f) { const double s = std::sqrt(f.a); return chain(f, s, 1.0 / (2.0 * s)); }
template <int N> inline Jet<N> sin_(const Jet<N>& f) { return chain(f, std::sin(f.a), std::cos(f.a)); }
template <int N> inline Jet<N> cos_(const Jet<N>& f) { return chain(f, std::cos(f.a), -std::sin(f.a)); }
template <int N> inline Jet<N> tan_(const Jet<N>& f) { const double t = std::tan(f.a); return chain(f, t, 1.0 + t * t); }
template <int N> inline Jet<N> atan_(const Jet<N>& f) { return chain(f, std::atan(f.a), 1.0 / (1.0 + f.a * f.a)); }
// ceres/jet.h abs: derivative is copysign(1, a)
template <int N> inline Jet<N> abs_(const Jet<N>& f) { return chain(f, std::fabs(f.a), std::copysign(1.0, f.a)); }
template <int N> inline Jet<N> atan2_(const Jet<N>& g, const Jet<N>& f) {
  // atan2(g, f): d = (f dg - g df) / (f^2 + g^2)
  Jet<N> h; const double t = 1.0 / (f.a * f.a + g.a * g.a); h.a = std::atan2(g.a, f.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]); return h; }
template <int N> inline double scalar(const Jet<N>& f) { return f.a; }

}  // namespace oracle
#endif  // ORACLE_JET_H_
