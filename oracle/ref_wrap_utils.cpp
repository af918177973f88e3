// ref_wrap_utils.cpp -- a C entry point around the REFERENCE's own mu::composePieceWisePol (TEST INFRASTRUCTURE).
//
// oracle/Makefile (target _ref) compiles neptune/src/utils.cpp where it lies under /root/reference, unmodified, against the
// Eigen stand-in and the field-only ROS message stand-ins of oracle/ref_stubs.  Records use the layout of
// include/neptune_b200.h: [0] pieces n, [1 .. 17] knot times, then x / y / z coefficients [16][4] each.
#include <vector>
#include "utils.hpp"

#define ORC_REC_TP 16 /* pieces a record can hold (oracle/neptune_oracle.c, NB_REC_TP of include/neptune_b200.h) */

static mt::PieceWisePol from_record(const double* r)
{
  mt::PieceWisePol p;
  const int n = (int)r[0];
  for (int i = 0; i <= n; i++) p.times.push_back(r[1 + i]);
  const double* c = r + 1 + (ORC_REC_TP + 1);
  for (int i = 0; i < n; i++)
  {
    p.coeff_x.push_back(Eigen::Matrix<double, 4, 1>(c[4 * i], c[4 * i + 1], c[4 * i + 2], c[4 * i + 3]));
    p.coeff_y.push_back(Eigen::Matrix<double, 4, 1>(c[ORC_REC_TP * 4 + 4 * i], c[ORC_REC_TP * 4 + 4 * i + 1], c[ORC_REC_TP * 4 + 4 * i + 2],
                                                    c[ORC_REC_TP * 4 + 4 * i + 3]));
    p.coeff_z.push_back(Eigen::Matrix<double, 4, 1>(c[2 * ORC_REC_TP * 4 + 4 * i], c[2 * ORC_REC_TP * 4 + 4 * i + 1],
                                                    c[2 * ORC_REC_TP * 4 + 4 * i + 2], c[2 * ORC_REC_TP * 4 + 4 * i + 3]));
  }
  return p;
}

// out_times [64], out_coeff [3][64][4]: the composed trajectory as the reference returns it (any number of pieces up to 64);
// t1 / t2 receive the knot times of p1 / p2 after the call (the reference adjusts times.front() of its arguments).
// Returns the number of pieces, 0 for the empty "dummy" result.
extern "C" int ref_compose_records(double t, double dc, const double* p1, const double* p2, double* out_times, double* out_coeff,
                                   double* t1, double* t2)
{
  mt::PieceWisePol a = from_record(p1), b = from_record(p2);
  mt::PieceWisePol p = mu::composePieceWisePol(t, dc, a, b);
  for (size_t i = 0; i < a.times.size(); i++) t1[i] = a.times[i];
  for (size_t i = 0; i < b.times.size(); i++) t2[i] = b.times[i];
  const int n = (int)p.coeff_x.size();
  for (size_t i = 0; i < p.times.size() && i < 64; i++) out_times[i] = p.times[i];
  for (int i = 0; i < n && i < 64; i++)
    for (int k = 0; k < 4; k++)
      out_coeff[(0 * 64 + i) * 4 + k] = p.coeff_x[i](k), out_coeff[(1 * 64 + i) * 4 + k] = p.coeff_y[i](k), out_coeff[(2 * 64 + i) * 4 + k] = p.coeff_z[i](k);
  return n;
}

// separator::Separator of the reference's own separator_glpk.cpp (submodules/separator), 2-D variants, with the LP engine
// supplied by the test through ref_stubs/glpk.h.  variant 0: solveModel(A, B) (:500-604); 1: solveModel(n, A, B)
// (:248-373); 2: solveModel(n, A, Aplus, B) (:375-498).  Points are [n][2].
#include "separator.hpp"
static Eigen::Matrix<double, 2, Eigen::Dynamic> points2(const double* p, int n)
{
  Eigen::Matrix<double, 2, Eigen::Dynamic> m(2, n);
  for (int i = 0; i < n; i++) m(0, i) = p[2 * i], m(1, i) = p[2 * i + 1];
  return m;
}
extern "C" int ref_separator_solve(int variant, const double* A, int nA, const double* Aplus, int nAp, const double* B, int nB, double* n_out)
{
  separator::Separator sep;
  Eigen::Vector3d n(0.0, 0.0, 0.0);
  bool ok;
  if (variant == 0)
    ok = sep.solveModel(points2(A, nA), points2(B, nB));
  else if (variant == 1)
    ok = sep.solveModel(n, points2(A, nA), points2(B, nB));
  else
    ok = sep.solveModel(n, points2(A, nA), points2(Aplus, nAp), points2(B, nB));
  n_out[0] = n(0), n_out[1] = n(1), n_out[2] = n(2);
  return ok ? 1 : 0;
}

// The 3-D entry point the reference's own test_separator.cpp drives (solveModel(n, d, pointsA, pointsB), :53-62 -> :64-246).
extern "C" int ref_separator_solve3d(const double* A, int nA, const double* B, int nB, double* n_out /*[4]: n, d*/)
{
  separator::Separator sep;
  std::vector<Eigen::Vector3d> a, b;
  for (int i = 0; i < nA; i++) a.push_back(Eigen::Vector3d(A[3 * i], A[3 * i + 1], A[3 * i + 2]));
  for (int i = 0; i < nB; i++) b.push_back(Eigen::Vector3d(B[3 * i], B[3 * i + 1], B[3 * i + 2]));
  Eigen::Vector3d n(0.0, 0.0, 0.0);
  double d = 0.0;
  const bool ok = sep.solveModel(n, d, a, b);
  n_out[0] = n(0), n_out[1] = n(1), n_out[2] = n(2), n_out[3] = d;
  return ok ? 1 : 0;
}
