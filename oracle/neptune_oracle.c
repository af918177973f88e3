/*
 * neptune_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See neptune_oracle.h for scope, usage rules and parity status ("parity unpinned":
 * no Gurobi/GLPK/CGAL here; pinned by HiGHS + known answers in tests/).
 *
 * Plain C restatement of the reference algorithm, one function per reference
 * function, each citing the file:line it follows (relative to /root/reference).
 * Dependency-free FP64.
 */
#include "neptune_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small dense helpers                                                        */
/* ------------------------------------------------------------------------- */

/* Gauss-Jordan inverse with partial pivoting (stands in for Eigen's .inverse(),
 * solver_gurobi_poly.cpp:93-94, neptune.cpp:64-65). */
static int mat_inverse(const double* A, int n, double* Ainv)
{
  double w[8 * 16];
  if (n > 8) return -1;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      w[i * 2 * n + j] = A[i * n + j];
      w[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; c++)
  {
    int p = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(w[r * 2 * n + c]) > fabs(w[p * 2 * n + c])) p = r;
    if (w[p * 2 * n + c] == 0.0) return -1;
    if (p != c)
      for (int j = 0; j < 2 * n; j++)
      {
        double t = w[c * 2 * n + j];
        w[c * 2 * n + j] = w[p * 2 * n + j];
        w[p * 2 * n + j] = t;
      }
    double piv = 1.0 / w[c * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) w[c * 2 * n + j] *= piv;
    for (int r = 0; r < n; r++)
      if (r != c)
      {
        double f = w[r * 2 * n + c];
        if (f != 0.0)
          for (int j = 0; j < 2 * n; j++) w[r * 2 * n + j] -= f * w[c * 2 * n + j];
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ainv[i * n + j] = w[i * 2 * n + n + j];
  return 0;
}

/* in-place lower Cholesky of a dense n x n (row-major, leading dim ld).  A pivot that cancels to
 * below 1e-13 of its original diagonal (rank deficiency) is floored there. */
static void chol_factor(double* A, int n, int ld)
{
  for (int j = 0; j < n; j++)
  {
    const double orig = A[j * ld + j];
    double d = orig;
    for (int k = 0; k < j; k++) d -= A[j * ld + k] * A[j * ld + k];
    const double floor_piv = 1e-13 * (orig > 0 ? orig : 1.0);
    if (!(d > floor_piv)) d = floor_piv;
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; i++)
    {
      double v = A[i * ld + j];
      for (int k = 0; k < j; k++) v -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = v / d;
    }
  }
}

static void chol_solve(const double* L, int n, int ld, double* b)
{
  for (int i = 0; i < n; i++)
  {
    double v = b[i];
    for (int k = 0; k < i; k++) v -= L[i * ld + k] * b[k];
    b[i] = v / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; i--)
  {
    double v = b[i];
    for (int k = i + 1; k < n; k++) v -= L[k * ld + i] * b[k];
    b[i] = v / L[i * ld + i];
  }
}

/* ------------------------------------------------------------------------- */
/* constants                                                                  */
/* ------------------------------------------------------------------------- */

/* MINVO position / velocity matrices for t in [0,1]: mader_types.hpp:152-162 */
static const double A_POS_MV[16] = {
  -3.4416308968564117698463178385282, 6.9895481477801393310755884158425, -4.4622887507045296828778191411402,
  0.91437149978080234369315348885721, 6.6792587327074839365081970754545, -11.845989901556746914934592496138,
  5.2523596690684613008670567069203, 0.0, -6.6792587327074839365081970754545, 8.1917862965657040064115790301003,
  -1.5981560640774179482548333908198, 0.085628500219197656306846511142794, 3.4416308968564117698463178385282,
  -3.3353445427890959784633650997421, 0.80808514571348655231020075007109,
  -0.0000000000000000084567769453869345852581318467855 };
static const double A_VEL_MV[9] = { 1.50000000000000, -2.36602540378444,  0.933012701892219, -3.0, 3.0, 0.0,
                                    1.50000000000000, -0.633974596215561, 0.0669872981077807 };

/* solver_gurobi_poly.cpp:35-62, :93-97 (A_rest_pos_basis_inverse_, A_rest_vel_basis_inverse321_);
 * neptune.cpp:64 (A_rest_pos_basis_inverse_ on [0,1]) */
void orc_basis(double T, double Ainv[16], double V[9], double Ainv01[16])
{
  double A[16], Av[9];
  const double cp[4] = { 1.0 / (T * T * T), 1.0 / (T * T), 1.0 / T, 1.0 };
  const double cv[3] = { 1.0 / (T * T), 1.0 / T, 1.0 };
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) A[i * 4 + j] = A_POS_MV[i * 4 + j] * cp[j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Av[i * 3 + j] = A_VEL_MV[i * 3 + j] * cv[j];
  mat_inverse(A, 4, Ainv);
  mat_inverse(Av, 3, V);
  for (int j = 0; j < 3; j++)
  {
    V[0 * 3 + j] *= 3.0;
    V[1 * 3 + j] *= 2.0;
  }
  if (Ainv01) mat_inverse(A_POS_MV, 4, Ainv01);
}

/* ------------------------------------------------------------------------- */
/* separator                                                                  */
/* ------------------------------------------------------------------------- */

#define SEP_EPS 1e-6 /* as the product (nb_common.cuh NB_SEP_EPS): see the note in nb_sep.cuh */

static int sep_check(const double* A, int nA, const double* B, int nB, const double l[3])
{
  for (int i = 0; i < nA; i++)
    if (!(l[0] * A[2 * i] + l[1] * A[2 * i + 1] + l[2] >= 1.0 - SEP_EPS)) return 0;
  for (int i = 0; i < nB; i++)
    if (!(l[0] * B[2 * i] + l[1] * B[2 * i + 1] + l[2] <= -1.0 + SEP_EPS)) return 0;
  return 1;
}

/* line with value +1 at pa (A side) and -1 at pb, normal along pa-pb */
static int sep_line_from_pair(const double pa[2], const double pb[2], double l[3], double* delta)
{
  double ux = pa[0] - pb[0], uy = pa[1] - pb[1];
  double d2 = ux * ux + uy * uy;
  if (!(d2 > 1e-24)) return 0;
  double d = sqrt(d2);
  *delta = d;
  l[0] = 2.0 * ux / d2;
  l[1] = 2.0 * uy / d2;
  l[2] = -(l[0] * (pa[0] + pb[0]) + l[1] * (pa[1] + pb[1])) * 0.5;
  return 1;
}

/* foot of the perpendicular from p on the infinite line (u,v) */
static int foot_on_line(const double p[2], const double u[2], const double v[2], double f[2])
{
  double ex = v[0] - u[0], ey = v[1] - u[1];
  double e2 = ex * ex + ey * ey;
  if (!(e2 > 1e-24)) return 0;
  double t = ((p[0] - u[0]) * ex + (p[1] - u[1]) * ey) / e2;
  f[0] = u[0] + t * ex;
  f[1] = u[1] + t * ey;
  return 1;
}

/*
 * separator::Separator::solveModel, 2-D (separator_glpk.cpp:248-373; the 4-arg overload
 * :375-498 is the same LP with A := A u A+, concatenated by the caller).
 * LP: find (n0,n1,d) with n.a+d >= 1 for a in A, n.b+d <= -1 for b in B, zero objective
 * (:273-285, :291-304).  GLPK returns an implementation-defined feasible vertex, which is
 * not reproducible without GLPK 4.65; the oracle (and the product) return the CANONICAL
 * feasible point instead: the minimum-norm (= maximum-margin) solution, which is unique.
 * It is found here by exhaustive enumeration of the possible support sets
 * (vertex/vertex, vertex/edge, edge/vertex) -- deliberately a different algorithm from the
 * product's closest-pair reduction.  Returns 1 (GLP_OPT/GLP_FEAS, :367) iff the LP is feasible.
 */
int orc_separate(const double* A, int nA, const double* B, int nB, double out[3])
{
  double best = -1.0, l[3], dl;
  int found = 0;
  out[0] = out[1] = out[2] = 0.0;
  /* vertex / vertex */
  for (int i = 0; i < nA; i++)
    for (int j = 0; j < nB; j++)
      if (sep_line_from_pair(A + 2 * i, B + 2 * j, l, &dl) && dl > best && sep_check(A, nA, B, nB, l))
      {
        best = dl;
        found = 1;
        memcpy(out, l, sizeof(l));
      }
  /* vertex of A / edge of B */
  for (int i = 0; i < nA; i++)
    for (int j = 0; j < nB; j++)
      for (int k = j + 1; k < nB; k++)
      {
        double f[2];
        if (!foot_on_line(A + 2 * i, B + 2 * j, B + 2 * k, f)) continue;
        if (sep_line_from_pair(A + 2 * i, f, l, &dl) && dl > best && sep_check(A, nA, B, nB, l))
        {
          best = dl;
          found = 1;
          memcpy(out, l, sizeof(l));
        }
      }
  /* edge of A / vertex of B */
  for (int i = 0; i < nB; i++)
    for (int j = 0; j < nA; j++)
      for (int k = j + 1; k < nA; k++)
      {
        double f[2];
        if (!foot_on_line(B + 2 * i, A + 2 * j, A + 2 * k, f)) continue;
        if (sep_line_from_pair(f, B + 2 * i, l, &dl) && dl > best && sep_check(A, nA, B, nB, l))
        {
          best = dl;
          found = 1;
          memcpy(out, l, sizeof(l));
        }
      }
  return found;
}

/*
 * The same rows as separator_glpk.cpp:99-110 (3-D) / :273-285 (2-D) handed to a generic
 * dense phase-1 simplex (Bland's rule): a literal "is this LP feasible" second opinion,
 * used to pin the solved flag and the known answer of test_separator.cpp:23-32.
 */
int orc_lp_separable(const double* A, int nA, const double* B, int nB, int dim)
{
  const int m = nA + nB, nv = dim + 1;
  const int ncol = 2 * nv + m + m; /* x+ , x-, surplus, artificial */
  if (m <= 0 || m > 128) return -1;
  double* T = (double*)calloc((size_t)(m + 1) * (ncol + 1), sizeof(double));
  int* basis = (int*)malloc(sizeof(int) * m);
  const int W = ncol + 1;
  for (int r = 0; r < m; r++)
  {
    const double* p = (r < nA) ? (A + dim * r) : (B + dim * (r - nA));
    double sgn = (r < nA) ? 1.0 : -1.0; /* B rows: -(n.b+d) >= 1 */
    for (int c = 0; c < dim; c++)
    {
      T[r * W + c] = sgn * p[c];
      T[r * W + nv + c] = -sgn * p[c];
    }
    T[r * W + dim] = sgn;
    T[r * W + nv + dim] = -sgn;
    T[r * W + 2 * nv + r] = -1.0;    /* surplus */
    T[r * W + 2 * nv + m + r] = 1.0; /* artificial */
    T[r * W + ncol] = 1.0;
    basis[r] = 2 * nv + m + r;
  }
  /* phase-1 objective: minimise sum of artificials -> reduced costs */
  for (int c = 0; c <= ncol; c++)
  {
    double s = 0.0;
    for (int r = 0; r < m; r++) s += T[r * W + c];
    T[m * W + c] = -s;
  }
  for (int r = 0; r < m; r++) T[m * W + 2 * nv + m + r] = 0.0;
  for (int iter = 0; iter < 10000; iter++)
  {
    int pc = -1;
    for (int c = 0; c < ncol; c++)
      if (T[m * W + c] < -1e-11)
      {
        pc = c;
        break;
      }
    if (pc < 0) break;
    int pr = -1;
    double bestr = 0.0;
    for (int r = 0; r < m; r++)
      if (T[r * W + pc] > 1e-11)
      {
        double ratio = T[r * W + ncol] / T[r * W + pc];
        if (pr < 0 || ratio < bestr - 1e-14 || (fabs(ratio - bestr) <= 1e-14 && basis[r] < basis[pr]))
        {
          pr = r;
          bestr = ratio;
        }
      }
    if (pr < 0) break;
    double piv = T[pr * W + pc];
    for (int c = 0; c <= ncol; c++) T[pr * W + c] /= piv;
    for (int r = 0; r <= m; r++)
      if (r != pr)
      {
        double f = T[r * W + pc];
        if (f != 0.0)
          for (int c = 0; c <= ncol; c++) T[r * W + c] -= f * T[pr * W + c];
      }
    basis[pr] = pc;
  }
  double obj = -T[m * W + ncol];
  free(T);
  free(basis);
  return obj < 1e-8;
}

/* ------------------------------------------------------------------------- */
/* convex hull, hull generation, sampling                                     */
/* ------------------------------------------------------------------------- */

static int cmp_pt(const void* a, const void* b)
{
  const double* p = (const double*)a;
  const double* q = (const double*)b;
  if (p[0] < q[0]) return -1;
  if (p[0] > q[0]) return 1;
  if (p[1] < q[1]) return -1;
  if (p[1] > q[1]) return 1;
  return 0;
}

static double cross3(const double* o, const double* a, const double* b)
{
  return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0]);
}

/*
 * cu::convexHullOfPoints2d (cgal_utils.cpp:157-174) -> CGAL::convex_hull_2.
 * CGAL 4.14.2 is not available; adopted convention (hard part 4):
 * strict extreme points, counter-clockwise, starting at the lexicographically smallest
 * point.  Andrew's monotone chain.  Returns the number of vertices written to out.
 */
int orc_convex_hull_2d(const double* pts, int n, double* out)
{
  if (n <= 0) return 0;
  double* p = (double*)malloc(sizeof(double) * 2 * n);
  double* h = (double*)malloc(sizeof(double) * 2 * (2 * n + 2));
  memcpy(p, pts, sizeof(double) * 2 * n);
  qsort(p, n, 2 * sizeof(double), cmp_pt);
  int m = 0;
  for (int i = 0; i < n; i++) /* unique */
    if (m == 0 || p[2 * i] != p[2 * (m - 1)] || p[2 * i + 1] != p[2 * (m - 1) + 1])
    {
      p[2 * m] = p[2 * i];
      p[2 * m + 1] = p[2 * i + 1];
      m++;
    }
  int k = 0;
  if (m <= 2)
  {
    memcpy(out, p, sizeof(double) * 2 * m);
    free(p);
    free(h);
    return m;
  }
  for (int i = 0; i < m; i++)
  {
    while (k >= 2 && cross3(h + 2 * (k - 2), h + 2 * (k - 1), p + 2 * i) <= 0) k--;
    h[2 * k] = p[2 * i];
    h[2 * k + 1] = p[2 * i + 1];
    k++;
  }
  for (int i = m - 2, t = k + 1; i >= 0; i--)
  {
    while (k >= t && cross3(h + 2 * (k - 2), h + 2 * (k - 1), p + 2 * i) <= 0) k--;
    h[2 * k] = p[2 * i];
    h[2 * k + 1] = p[2 * i + 1];
    k++;
  }
  k--; /* last point equals the first */
  memcpy(out, h, sizeof(double) * 2 * k);
  free(p);
  free(h);
  return k;
}

static int lower_bound_d(const double* a, int n, double v)
{
  int i = 0;
  while (i < n && a[i] < v) i++;
  return i;
}
static int upper_bound_d(const double* a, int n, double v)
{
  int i = 0;
  while (i < n && !(v < a[i])) i++;
  return i;
}
static int sat_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * Neptune::convexHullOfInterval2d + vertexesOfInterval2d (neptune.cpp:288-309, :349-452):
 * MINVO control points of every piece of (times, cx, cy) overlapping [t_start, t_end],
 * with partial-interval time scaling (:400-429), inflated to 4 box corners (:442-445) and
 * hulled; also the un-inflated hull (polygon2).  idx = {index_first_interval, index_last_interval}
 * after saturation (:382-389).
 */
void orc_hull_of_interval(const double* times, int nt, const double* cx, const double* cy, double t_start,
                          double t_end, double T_span, const double delta[3], double* hull, int* hull_n,
                          double* hull2, int* hull2_n, int idx[2])
{
  double Ainv[16], V[9], Ainv01[16];
  orc_basis(T_span, Ainv, V, Ainv01);
  const int np = nt - 1;
  int first = lower_bound_d(times, nt, t_start) - 1;
  int last = upper_bound_d(times, nt, t_end) - 1;
  first = sat_i(first, 0, np - 1);
  last = sat_i(last, 0, np - 1);
  if (idx)
  {
    idx[0] = first;
    idx[1] = last;
  }
  double pts[2 * 4 * 4 * (ORC_NPOL_MAX * 2 + 2)], pts2[2 * 4 * (ORC_NPOL_MAX * 2 + 2)];
  int npts = 0, npts2 = 0;
  const double dnorm = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  for (int i = first; i <= last && npts2 < 4 * (ORC_NPOL_MAX * 2 + 1); i++)
  {
    double t;
    if (i != last)
      t = times[i + 1] - times[i];
    else if (t_end > times[i + 1])
      t = times[i + 1] - times[i];
    else
      t = t_end - times[i];
    if (t > T_span)
      t = T_span;
    else if (t < 0)
      t = 0;
    const double sc[4] = { t * t * t, t * t, t, 1.0 };
    for (int k = 0; k < 4; k++)
    {
      double x = 0, y = 0;
      for (int r = 0; r < 4; r++)
      {
        x += cx[4 * i + r] * sc[r] * Ainv01[r * 4 + k];
        y += cy[4 * i + r] * sc[r] * Ainv01[r * 4 + k];
      }
      if (dnorm < 1e-6)
      {
        pts[2 * npts] = x;
        pts[2 * npts + 1] = y;
        npts++;
      }
      else
      {
        const double sx[4] = { 1, 1, -1, -1 }, sy[4] = { 1, -1, -1, 1 };
        for (int c = 0; c < 4; c++)
        {
          pts[2 * npts] = x + sx[c] * delta[0];
          pts[2 * npts + 1] = y + sy[c] * delta[1];
          npts++;
        }
      }
      pts2[2 * npts2] = x;
      pts2[2 * npts2 + 1] = y;
      npts2++;
    }
  }
  *hull_n = orc_convex_hull_2d(pts, npts, hull);
  if (hull2) *hull2_n = orc_convex_hull_2d(pts2, npts2, hull2);
}

/*
 * Neptune::SamplePointsOfIntervals (neptune.cpp:500-566): S+1 samples per interval of one
 * other agent, out[num_pol][S+1][2]; idx_out[num_pol][S+1] = index_interval (:519-520, or
 * the "outside" branch index :545).
 */
void orc_sample_interval_points(const double* times, int nt, const double* cx, const double* cy,
                                double t_start, double t_end, int num_pol, int S, double* out, int* idx_out)
{
  const double deltaT = (t_end - t_start) / (1.0 * num_pol);
  const int np = nt - 1;
  for (int i = 0; i < num_pol; i++)
    for (int j = 0; j <= S; j++)
    {
      double ts = t_start + deltaT * i + deltaT / S * j;
      int low = upper_bound_d(times, nt, ts);
      int ii;
      double te;
      if (low != nt)
      {
        ii = sat_i(low - 1, 0, np - 1);
        te = ts - times[ii];
        if (te < 0)
          te = 0;
        else if (te > deltaT)
          te = deltaT;
      }
      else
      {
        int k = low - 1; /* last time index */
        te = times[k] - times[k - 1];
        ii = k - 1;
      }
      const double tv[4] = { te * te * te, te * te, te, 1.0 };
      double x = 0, y = 0;
      for (int r = 0; r < 4; r++)
      {
        x += cx[4 * ii + r] * tv[r];
        y += cy[4 * ii + r] * tv[r];
      }
      out[(i * (S + 1) + j) * 2] = x;
      out[(i * (S + 1) + j) * 2 + 1] = y;
      if (idx_out) idx_out[i * (S + 1) + j] = ii;
    }
}

/* ------------------------------------------------------------------------- */
/* GJK                                                                        */
/* ------------------------------------------------------------------------- */

static int gjk_furthest(const double* v, int n, double dx, double dy)
{
  double best = dx * v[0] + dy * v[1];
  int idx = 0;
  for (int i = 1; i < n; i++)
  {
    double p = dx * v[2 * i] + dy * v[2 * i + 1];
    if (p > best)
    {
      best = p;
      idx = i;
    }
  }
  return idx;
}
static void gjk_support(const double* v1, int n1, const double* v2, int n2, double dx, double dy, double s[2])
{
  int i = gjk_furthest(v1, n1, dx, dy), j = gjk_furthest(v2, n2, -dx, -dy);
  s[0] = v1[2 * i] - v2[2 * j];
  s[1] = v1[2 * i + 1] - v2[2 * j + 1];
}
/* b*(a.c) - a*(b.c)  (gjk.cpp:25-28) */
static void gjk_triple(const double a[2], const double b[2], const double c[2], double r[2])
{
  double ac = a[0] * c[0] + a[1] * c[1], bc = b[0] * c[0] + b[1] * c[1];
  r[0] = b[0] * ac - a[0] * bc;
  r[1] = b[1] * ac - a[1] * bc;
}

/* gjk::collision (gjk.cpp:76-148): yes/no GJK in 2-D on vertex arrays */
int orc_gjk_collision(const double* v1, int n1, const double* v2, int n2)
{
  double simplex[3][2], a[2], b[2], c[2], d[2], ao[2], ab[2], ac[2], abp[2], acp[2];
  double p1[2] = { 0, 0 }, p2[2] = { 0, 0 };
  int index = 0;
  for (int i = 0; i < n1; i++)
  {
    p1[0] += v1[2 * i];
    p1[1] += v1[2 * i + 1];
  }
  for (int i = 0; i < n2; i++)
  {
    p2[0] += v2[2 * i];
    p2[1] += v2[2 * i + 1];
  }
  d[0] = p1[0] / n1 - p2[0] / n2;
  d[1] = p1[1] / n1 - p2[1] / n2;
  if (d[0] == 0 && d[1] == 0) d[0] = 1.0;
  gjk_support(v1, n1, v2, n2, d[0], d[1], simplex[0]);
  a[0] = simplex[0][0];
  a[1] = simplex[0][1];
  if (a[0] * d[0] + a[1] * d[1] <= 0) return 0;
  d[0] = -a[0];
  d[1] = -a[1];
  for (int guard = 0; guard < 1000; guard++)
  {
    ++index;
    gjk_support(v1, n1, v2, n2, d[0], d[1], simplex[index]);
    a[0] = simplex[index][0];
    a[1] = simplex[index][1];
    if (a[0] * d[0] + a[1] * d[1] <= 0) return 0;
    ao[0] = -a[0];
    ao[1] = -a[1];
    if (index < 2)
    {
      b[0] = simplex[0][0];
      b[1] = simplex[0][1];
      ab[0] = b[0] - a[0];
      ab[1] = b[1] - a[1];
      gjk_triple(ab, ao, ab, d);
      if (sqrt(d[0] * d[0] + d[1] * d[1]) == 0)
      {
        d[0] = ab[1];
        d[1] = -ab[0];
      }
      continue;
    }
    b[0] = simplex[1][0];
    b[1] = simplex[1][1];
    c[0] = simplex[0][0];
    c[1] = simplex[0][1];
    ab[0] = b[0] - a[0];
    ab[1] = b[1] - a[1];
    ac[0] = c[0] - a[0];
    ac[1] = c[1] - a[1];
    gjk_triple(ab, ac, ac, acp);
    if (acp[0] * ao[0] + acp[1] * ao[1] >= 0)
    {
      d[0] = acp[0];
      d[1] = acp[1];
    }
    else
    {
      gjk_triple(ac, ab, ab, abp);
      if (abp[0] * ao[0] + abp[1] * ao[1] < 0) return 1;
      simplex[0][0] = simplex[1][0];
      simplex[0][1] = simplex[1][1];
      d[0] = abp[0];
      d[1] = abp[1];
    }
    simplex[1][0] = simplex[2][0];
    simplex[1][1] = simplex[2][1];
    --index;
  }
  return 0;
}

/*
 * Neptune::trajsAndPwpAreInCollision2d (neptune.cpp:767-806): my optimised pwp (n intervals of
 * length T_span from t_start) against another agent's trajectory; per interval the GJK test of my
 * MINVO control points (P * A_rest_pos_basis_t_inverse_, :789) against the other's inflated hull.
 */
int orc_pwp_collides(const double* coeff, int n, double t_start, double T_span, const double* times, int nt,
                     const double* cx, const double* cy, const double delta[3])
{
  double Ainv[16], V[9];
  orc_basis(T_span, Ainv, V, 0);
  const double t_end = t_start + T_span * n;
  const double deltaT = (t_end - t_start) / n;
  if (fabs(deltaT - T_span) > 0.1) return 1; /* :772-780 */
  for (int i = 0; i < n; i++)
  {
    double A[8], hull[2 * ORC_HMAX];
    int hn, idx[2];
    for (int k = 0; k < 4; k++)
    {
      double x = 0, y = 0;
      for (int r = 0; r < 4; r++)
      {
        x += coeff[4 * i + r] * Ainv[r * 4 + k];
        y += coeff[32 + 4 * i + r] * Ainv[r * 4 + k];
      }
      A[2 * k] = x;
      A[2 * k + 1] = y;
    }
    orc_hull_of_interval(times, nt, cx, cy, t_start + deltaT * i, t_start + deltaT * (i + 1), T_span, delta, hull, &hn,
                         0, 0, idx);
    if (orc_gjk_collision(hull, hn, A, 4)) return 1;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* entanglement chain                                                         */
/* ------------------------------------------------------------------------- */

/* eu::vectorWedge2 (entangle_utils.cpp:16-27): (b-a) x (c-a); optionally returns ab, ac */
static double wedge2(const double a[2], const double b[2], const double c[2], double ab[2], double ac[2])
{
  double abx = b[0] - a[0], aby = b[1] - a[1], acx = c[0] - a[0], acy = c[1] - a[1];
  if (ab)
  {
    ab[0] = abx;
    ab[1] = aby;
    ac[0] = acx;
    ac[1] = acy;
  }
  return abx * acy - acx * aby;
}

/* classification ratio on the dominant coordinate (entangle_utils.cpp:1164-1172, :1194-1202);
 * unqualified abs(double) there is fabs on GCC >= 6 (hard part 3) */
static double cross_ratio(const double ab[2], const double ac[2])
{
  if (fabs(ab[1] * ac[1]) > fabs(ab[0] * ac[0])) return ab[1] / ac[1];
  return ab[0] / ac[0];
}

/* eu::entangleHSigToAddAgentInd, 8-arg (entangle_utils.cpp:1129-1228). toadd holds (id,case)
 * pairs; returns the new count. */
int orc_hsig_agent(int* toadd, int nadd, const double pk[2], const double pk1[2], const double pik[2],
                   const double pik1[2], const double pb_self[2], const double* bend, int nbend, int agent_id)
{
  int base_add = 0;
  for (int i = 0; i < nbend; i++)
  {
    double ab[2], ac[2], c1, c2;
    const double* bi = bend + 2 * i;
    const int last = (i == nbend - 1);
    if (!last)
    {
      c1 = wedge2(pk, bend + 2 * (i + 1), bi, ab, ac);
      c2 = wedge2(pk1, bend + 2 * (i + 1), bi, 0, 0);
    }
    else
    {
      c1 = wedge2(pk, pik, bi, ab, ac);
      c2 = wedge2(pk1, pik1, bi, 0, 0);
    }
    if (last)
    {
      double fb[2], fc[2];
      double f1 = wedge2(pb_self, pik, bi, fb, fc);
      double f2 = wedge2(pb_self, pik1, bi, 0, 0);
      if (f1 * f2 < 0)
      {
        double a = cross_ratio(fb, fc);
        if (a < 0)
        {
        }
        else if (a < 1)
        {
          toadd[2 * nadd] = agent_id;
          toadd[2 * nadd + 1] = 1;
          nadd++;
        }
        else if (i == 0)
        {
          toadd[2 * nadd] = agent_id;
          toadd[2 * nadd + 1] = 0;
          nadd++;
        }
        base_add = 1;
      }
    }
    if (c1 * c2 < 0)
    {
      double a = cross_ratio(ab, ac);
      if (a < 0)
      {
        toadd[2 * nadd] = agent_id;
        toadd[2 * nadd + 1] = i + 2;
        nadd++;
      }
      else if (a < 1 && last)
      {
        toadd[2 * nadd] = agent_id;
        toadd[2 * nadd + 1] = 1;
        nadd++;
      }
      else if (a >= 1 && i == 0)
      {
        toadd[2 * nadd] = agent_id;
        toadd[2 * nadd + 1] = 0;
        nadd++;
      }
    }
  }
  if (base_add && nadd >= 2 && toadd[2 * (nadd - 1)] == toadd[2 * (nadd - 2)] &&
      toadd[2 * (nadd - 1) + 1] == toadd[2 * (nadd - 2) + 1])
    nadd -= 2;
  return nadd;
}

/* eu::entangleHSigToAddAgentInd, 9-arg (entangle_utils.cpp:820-1127): the online tracker's form, aware of a
 * bend point of the other tether having been added or released since the last check.  *stop = k where the
 * reference would print "stop k" and exit(-1) (:917, :981, :1073, :1107). */
#define PUSH9(cs_)                   \
  {                                  \
    toadd[2 * nadd] = agent_id;      \
    toadd[2 * nadd + 1] = (cs_);     \
    nadd++;                          \
  }
int orc_hsig_agent9(int* toadd, int nadd, const double pk[2], const double pk1[2], const double pik[2],
                    const double pik1[2], const double pb_self[2], const double* bend, int nbend, const double* prev,
                    int nprev, int agent_id, int* stop)
{
  if (nprev == nbend) return orc_hsig_agent(toadd, nadd, pk, pk1, pik, pik1, pb_self, bend, nbend, agent_id); /* :826-830 */
  int base_add = 0;
  if (nbend == 0 || nprev == 0) return nadd; /* "bendpts empty!" :833-836 */
  const double* pback = prev + 2 * (nprev - 1);
  if (nbend < nprev)
  { /* released from a bend point (:837-986) */
    const double* pback2 = prev + 2 * (nprev - 2);
    for (int i = 0; i < nbend; i++)
    {
      double ab[2], ac[2], abp[2], acp[2], c1, c2, c1p = 0.0;
      const double* bi = bend + 2 * i;
      const int last = (i == nbend - 1);
      if (!last)
      {
        c1 = wedge2(pk, bend + 2 * (i + 1), bi, ab, ac);
        c2 = wedge2(pk1, bend + 2 * (i + 1), bi, 0, 0);
      }
      else
      {
        c1 = wedge2(pk, pik, pback, ab, ac);
        c2 = wedge2(pk1, pik1, bi, 0, 0);
        c1p = wedge2(pk, pback, pback2, abp, acp);
      }
      if (last)
      {
        double fb[2], fc[2];
        double f1 = wedge2(pb_self, pik, pback, 0, 0);
        double f2 = wedge2(pb_self, pik1, bi, fb, fc);
        double f1p = 0.0;
        if (i == 0) f1p = wedge2(pb_self, pback, pback2, 0, 0);
        if (f1 * f2 < 0)
        {
          double a = cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
            PUSH9(1)
          base_add = 1;
        }
        if (i == 0 && f1p * f2 < 0)
        {
          double a = cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
          {
          }
          else
          {
            PUSH9(0)
            *stop = 4;
          }
          base_add = 1;
        }
      }
      int added_inbtw = 0;
      if (c1 * c2 < 0)
      {
        double a = cross_ratio(ab, ac);
        if (a < 0)
        {
          PUSH9(i + 2)
          added_inbtw = 1;
        }
        else if (a < 1 && last)
          PUSH9(1)
      }
      if (last && c1p * c2 < 0)
      {
        double a = cross_ratio(abp, acp);
        if (a < 0 && !added_inbtw)
          PUSH9(i + 2)
        else if (a < 1 && last)
        {
        }
        else if (a >= 1 && i == 0)
        {
          PUSH9(0)
          *stop = 3;
        }
      }
    }
  }
  else
  { /* a bend point was added (:987-1116) */
    for (int i = 0; i < nbend; i++)
    {
      double ab[2], ac[2], c1, c2;
      const double* bi = bend + 2 * i;
      if (i == nbend - 1)
      {
        c1 = wedge2(pk, pik, pback, 0, 0);
        c2 = wedge2(pk1, pik1, bi, ab, ac);
      }
      else if (i == nbend - 2)
      {
        c1 = wedge2(pk, pik, pback, 0, 0);
        c2 = wedge2(pk1, bend + 2 * (i + 1), bi, ab, ac);
      }
      else
      {
        c1 = wedge2(pk, bend + 2 * (i + 1), bi, ab, ac);
        c2 = wedge2(pk1, bend + 2 * (i + 1), bi, 0, 0);
      }
      if (i == nbend - 1)
      {
        double fb[2], fc[2];
        double f1 = wedge2(pb_self, pik, pback, 0, 0);
        double f2 = wedge2(pb_self, pik1, bi, fb, fc);
        if (f1 * f2 < 0)
        {
          double a = cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
            PUSH9(1)
          base_add = 1;
        }
      }
      if (i == 0 && nbend == 2)
      {
        double fb[2], fc[2];
        double f1 = wedge2(pb_self, pik, pback, 0, 0);
        double f2 = wedge2(pb_self, bend + 2 * (i + 1), bi, fb, fc);
        if (f1 * f2 < 0)
        {
          double a = cross_ratio(fb, fc);
          if (a < 0)
          {
          }
          else if (a < 1)
          {
          }
          else
          {
            PUSH9(0)
            *stop = 2;
          }
          base_add = 1;
        }
      }
      if (c1 * c2 < 0)
      {
        double a = cross_ratio(ab, ac);
        if (a < 0)
          PUSH9(i + 2)
        else if (a < 1 && i == nbend - 1)
          PUSH9(1)
        else if (a >= 1 && i == 0)
        {
          PUSH9(0)
          *stop = 1;
        }
      }
    }
  }
  if (base_add && nadd >= 2 && toadd[2 * (nadd - 1)] == toadd[2 * (nadd - 2)] &&
      toadd[2 * (nadd - 1) + 1] == toadd[2 * (nadd - 2) + 1])
    nadd -= 2;
  return nadd;
}

/* NeptuneRos::updateEntStateStaticObs (neptune_ros.cpp:798-850): one odometry tick of the online tracker.
 * es = entangle_state_ (in/out); prev_pos [N+1][2] = previousCheckingPos_ and prev_pos_agent [N][2] =
 * previousCheckingPosAgent_ (both in/out); latest [N][2] = latestCheckingPosAgent_; elapsed_ms replaces the
 * wall-clock timer of :803-804.  Returns 0 updated, 1 skipped by the gate, -k "stop k" (the reference exits),
 * -100 storage overflow. */
int orc_track(orc_ent* es, const orc_ectx* cx, const int* bp_cnt_prev, const double* bp_xy_prev, double* prev_pos,
              double* prev_pos_agent, const double* latest, const double cur[2], double elapsed_ms)
{
  const int N = cx->N, tcap = 4 * (cx->N + cx->M) + 16;
  {
    const double dx = prev_pos[2 * N] - cur[0], dy = prev_pos[2 * N + 1] - cur[1];
    if (sqrt(dx * dx + dy * dy) < 0.05 && elapsed_ms < 100) return 1;
  }
  int* toadd = (int*)malloc(sizeof(int) * 2 * tcap);
  int nadd = 0, rc = 0, stop = 0;
  for (int i = 0; i < N && !rc; i++)
  {
    if (i == cx->self) continue;
    if (prev_pos_agent[2 * i] < -900 || cx->bp_cnt[i] == 0) continue; /* agent msg not received yet (:811) */
    if (nadd + cx->bp_cnt[i] + 3 > tcap)
    {
      rc = -100;
      break;
    }
    nadd = orc_hsig_agent9(toadd, nadd, prev_pos + 2 * i, cur, prev_pos_agent + 2 * i, latest + 2 * i, cx->pb + 2 * cx->self,
                           cx->bp_xy + 2 * cx->bp_max * i, cx->bp_cnt[i], bp_xy_prev + 2 * cx->bp_max * i, bp_cnt_prev[i],
                           i + 1, &stop);
    prev_pos[2 * i] = cur[0], prev_pos[2 * i + 1] = cur[1];
    prev_pos_agent[2 * i] = latest[2 * i], prev_pos_agent[2 * i + 1] = latest[2 * i + 1];
  }
  if (!rc && nadd + cx->M > tcap) rc = -100;
  if (!rc)
  {
    nadd = orc_hsig_static(toadd, nadd, prev_pos + 2 * N, cur, cx->strep, cx->M, N);
    if (orc_add_alpha_beta(toadd, nadd, es, prev_pos + 2 * N, cx))
      rc = -100;
    else
      orc_update_bend_pts(es, cur, cx);
    prev_pos[2 * N] = cur[0], prev_pos[2 * N + 1] = cur[1];
  }
  free(toadd);
  if (stop) return -stop;
  return rc;
}

/* eu::entangleHSigToAddStatic (entangle_utils.cpp:1231-1277) */
int orc_hsig_static(int* toadd, int nadd, const double pk[2], const double pk1[2], const double* strep,
                    int M, int N)
{
  for (int i = 0; i < M; i++)
  {
    const double* pbi = strep + 4 * i;     /* col(0) */
    const double* pik = strep + 4 * i + 2; /* col(1) */
    double ab[2], ac[2];
    double c1 = wedge2(pk, pik, pbi, ab, ac);
    double c2 = wedge2(pk1, pik, pbi, 0, 0);
    if (c1 * c2 < 0)
    {
      double a = cross_ratio(ab, ac);
      if (a < 0)
      {
      }
      else if (a < 1)
      {
        toadd[2 * nadd] = N + i + 1;
        toadd[2 * nadd + 1] = 1;
        nadd++;
      }
      else
      {
        toadd[2 * nadd] = N + i + 1;
        toadd[2 * nadd + 1] = 0;
        nadd++;
      }
    }
  }
  return nadd;
}

/* eu::getBendPt2d (entangle_utils.cpp:1649-1679) */
static void bend_pt(double bp[2], const orc_ent* es, const orc_ectx* cx)
{
  if (es->n_bend == 0)
  {
    bp[0] = cx->pb[2 * cx->self];
    bp[1] = cx->pb[2 * cx->self + 1];
    return;
  }
  int id = es->alpha[2 * es->bend[es->n_bend - 1]], cs = es->alpha[2 * es->bend[es->n_bend - 1] + 1];
  if (id <= cx->N && id >= 1)
  {
    bp[0] = cx->pb[2 * (id - 1)];
    bp[1] = cx->pb[2 * (id - 1) + 1];
  }
  else if (id > cx->N)
  {
    bp[0] = cx->strep[4 * (id - cx->N - 1) + 2 * cs];
    bp[1] = cx->strep[4 * (id - cx->N - 1) + 2 * cs + 1];
  }
}

/* eu::calculateBetaForCase (entangle_utils.cpp:1709-1722) */
static double beta_for_case(int id, int cs, const double pk[2], const double bp[2], const orc_ectx* cx)
{
  if (id <= cx->N) return 0.0;
  return wedge2(pk, cx->strep + 4 * (id - cx->N - 1) + 2 * cs, bp, 0, 0);
}

/* eu::breakcondition (entangle_utils.cpp:1608-1647) */
static int break_condition(const int add[2], const int inlist[2], int N, int idx_to_check, int idx_last_bend)
{
  if (add[0] <= N && add[1] >= 2)
  {
    if (idx_to_check <= idx_last_bend) return 1;
  }
  else if (add[0] <= N && add[1] < 2)
  {
  }
  else if (add[0] > N)
  {
    if (inlist[0] > N || idx_to_check <= idx_last_bend) return 1;
  }
  return 0;
}

/* eu::addAlphaBetaToList (entangle_utils.cpp:1402-1534). Consumes toadd; returns 0, or -1 when
 * the storage capacity cx->cap would be exceeded (loud failure instead of a silent drop). */
int orc_add_alpha_beta(int* toadd, int nadd, orc_ent* es, const double pk[2], const orc_ectx* cx)
{
  const int N = cx->N;
  int have = 1;
  while (have)
  {
    have = 0;
    const int b = es->n_bend == 0 ? -1 : es->bend[es->n_bend - 1];
    for (int i = 0; i < nadd && !have; i++)
    {
      const int aid = toadd[2 * i], acs = toadd[2 * i + 1];
      for (int j = es->n_alpha - 1; j >= 0; j--)
      {
        const int lid = es->alpha[2 * j], lcs = es->alpha[2 * j + 1];
        int nb = (aid <= N) ? cx->bp_cnt[aid - 1] : 0;
        int cond = (lid == aid && lcs == acs) ||
                   (aid <= N && lid == aid && acs >= nb + 1 && acs < lcs) ||
                   (aid <= N && lid == aid && lcs >= 2 && acs >= 2 && abs(acs - lcs) == 1 && j > b);
        if (cond)
        {
          es->active[aid - 1] -= 1;
          memmove(toadd + 2 * i, toadd + 2 * (i + 1), sizeof(int) * 2 * (nadd - i - 1));
          nadd--;
          memmove(es->alpha + 2 * j, es->alpha + 2 * (j + 1), sizeof(int) * 2 * (es->n_alpha - j - 1));
          memmove(es->beta + j, es->beta + j + 1, sizeof(double) * (es->n_alpha - j - 1));
          es->n_alpha--;
          if (j == b)
          {
            es->n_bend--;
            double bp[2];
            bend_pt(bp, es, cx);
            for (int k = j; k < es->n_alpha; k++)
              es->beta[k] = beta_for_case(es->alpha[2 * k], es->alpha[2 * k + 1], pk, bp, cx);
          }
          else if (j < b)
          {
            es->bend[es->n_bend - 1] = b - 1;
            for (int k = es->n_bend - 2; k >= 0; k--)
            {
              if (es->bend[k] > j)
                es->bend[k] -= 1;
              else
                break;
            }
          }
          have = 1;
          break;
        }
        const int a2[2] = { aid, acs }, l2[2] = { lid, lcs };
        if (break_condition(a2, l2, N, j, b)) break;
      }
    }
  }
  if (nadd == 0) return 0;
  double bp[2];
  bend_pt(bp, es, cx);
  for (int i = 0; i < nadd; i++)
  {
    if (es->n_alpha >= cx->cap) return -1;
    es->alpha[2 * es->n_alpha] = toadd[2 * i];
    es->alpha[2 * es->n_alpha + 1] = toadd[2 * i + 1];
    es->active[toadd[2 * i] - 1] += 1;
    es->beta[es->n_alpha] = beta_for_case(toadd[2 * i], toadd[2 * i + 1], pk, bp, cx);
    es->n_alpha++;
  }
  return 0;
}

static void bend_coord(double bp[2], int id, int cs, const orc_ectx* cx)
{
  if (id <= cx->N)
  {
    bp[0] = cx->pb[2 * (id - 1)];
    bp[1] = cx->pb[2 * (id - 1) + 1];
  }
  else
  {
    bp[0] = cx->strep[4 * (id - cx->N - 1) + 2 * cs];
    bp[1] = cx->strep[4 * (id - cx->N - 1) + 2 * cs + 1];
  }
}

/* eu::updateBendPts (entangle_utils.cpp:1536-1604) */
void orc_update_bend_pts(orc_ent* es, const double pk1[2], const orc_ectx* cx)
{
  double bp[2];
  bend_pt(bp, es, cx);
  int idx_new = -1;
  int start = es->n_bend == 0 ? -1 : es->bend[es->n_bend - 1];
  for (int i = start + 1; i < es->n_alpha; i++)
  {
    double beta = beta_for_case(es->alpha[2 * i], es->alpha[2 * i + 1], pk1, bp, cx);
    if (beta * es->beta[i] < -1e-7) idx_new = i;
  }
  if (idx_new > -1)
  {
    if (es->n_bend < cx->cap) es->bend[es->n_bend++] = idx_new;
    double nb[2];
    bend_coord(nb, es->alpha[2 * idx_new], es->alpha[2 * idx_new + 1], cx);
    for (int i = idx_new + 1; i < es->n_alpha; i++)
      es->beta[i] = beta_for_case(es->alpha[2 * i], es->alpha[2 * i + 1], pk1, nb, cx);
    return;
  }
  while (es->n_bend > 0)
  {
    double prev[2];
    if (es->n_bend == 1)
    {
      prev[0] = cx->pb[2 * cx->self];
      prev[1] = cx->pb[2 * cx->self + 1];
    }
    else
    {
      int q = es->bend[es->n_bend - 2];
      bend_coord(prev, es->alpha[2 * q], es->alpha[2 * q + 1], cx);
    }
    int lb = es->bend[es->n_bend - 1];
    double beta = beta_for_case(es->alpha[2 * lb], es->alpha[2 * lb + 1], pk1, prev, cx);
    if (beta * es->beta[lb] > 1e-7)
    {
      for (int k = lb + 1; k < es->n_alpha; k++)
        es->beta[k] = beta_for_case(es->alpha[2 * k], es->alpha[2 * k + 1], pk1, prev, cx);
      es->n_bend--;
    }
    else
      break;
  }
}

/* Neptune::PredictAlphasBetas (neptune.cpp:976-1008): es holds entangle_state_ on entry,
 * entangle_state_A on exit. samp0[j] = SampledPointsForAll[j][0].col(0); known[j]=0 <=> empty. */
int orc_predict(orc_ent* es, const orc_ectx* cx, const double* prev_pos, const double* prev_pos_agent,
                const double cur[2], const double* samp0, const unsigned char* known)
{
  const int tcap = 4 * (cx->N + cx->M) + 16;
  int* toadd = (int*)malloc(sizeof(int) * 2 * tcap);
  int nadd = 0, rc = 0;
  for (int i = 0; i < cx->N; i++)
  {
    if (i == cx->self || !known[i]) continue;
    if (nadd + cx->bp_cnt[i] + 2 > tcap)
    {
      rc = -1;
      break;
    }
    nadd = orc_hsig_agent(toadd, nadd, prev_pos + 2 * i, cur, prev_pos_agent + 2 * i, samp0 + 2 * i,
                          cx->pb + 2 * cx->self, cx->bp_xy + 2 * cx->bp_max * i, cx->bp_cnt[i], i + 1);
  }
  if (!rc && nadd + cx->M > tcap) rc = -1;
  if (!rc)
  {
    nadd = orc_hsig_static(toadd, nadd, prev_pos + 2 * cx->N, cur, cx->strep, cx->M, cx->N);
    rc = orc_add_alpha_beta(toadd, nadd, es, prev_pos + 2 * cx->N, cx);
    if (!rc) orc_update_bend_pts(es, cur, cx);
  }
  free(toadd);
  return rc;
}

static void eval_xy(const double* cxy, int n, int i, double t, double p[2])
{
  const double* x = cxy + 4 * i;
  const double* y = cxy + 4 * n + 4 * i;
  /* P * [t^3 t^2 t 1] as a plain dot product (kinodynamic_search.cpp:122-127) */
  const double tv[4] = { t * t * t, t * t, t, 1.0 };
  p[0] = x[0] * tv[0] + x[1] * tv[1] + x[2] * tv[2] + x[3] * tv[3];
  p[1] = y[0] * tv[0] + y[1] * tv[1] + y[2] * tv[2] + y[3] * tv[3];
}

/* one S-step pass over interval `ii` shared by the front-end chain (kinodynamic_search.cpp:813-881)
 * and the post-check (:909-972). limit = list-length bound. returns 1 entangled, 0 fine, -1 overflow */
static int ent_interval_pass(orc_ent* es, const orc_ectx* cx, int n, const double* cxy, int ii,
                             const double* samp, const unsigned char* known, int num_pol, int S, double T,
                             int limit, int* toadd, int tcap, int* act_old)
{
  const int NA = cx->N + cx->M;
  double pk[2] = { cxy[4 * ii + 3], cxy[4 * n + 4 * ii + 3] }, pk1[2];
  memcpy(act_old, es->active, sizeof(int) * NA);
  for (int j = 1; j <= S; j++)
  {
    int nadd = 0;
    double t = (j < S) ? T * j / S : T;
    eval_xy(cxy, n, ii, t, pk1);
    for (int a = 0; a < cx->N; a++)
    {
      if (a == cx->self || !known[a]) continue;
      const double *pik, *pik1;
      if (ii > num_pol - 1)
      {
        pik = samp + ((size_t)(a * num_pol + (num_pol - 1)) * (S + 1) + S) * 2;
        pik1 = pik;
      }
      else
      {
        pik = samp + ((size_t)(a * num_pol + ii) * (S + 1) + (j - 1)) * 2;
        pik1 = samp + ((size_t)(a * num_pol + ii) * (S + 1) + j) * 2;
      }
      if (nadd + cx->bp_cnt[a] + 2 > tcap) return -1;
      nadd = orc_hsig_agent(toadd, nadd, pk, pk1, pik, pik1, cx->pb + 2 * cx->self,
                            cx->bp_xy + 2 * cx->bp_max * a, cx->bp_cnt[a], a + 1);
    }
    if (nadd + cx->M > tcap) return -1;
    nadd = orc_hsig_static(toadd, nadd, pk, pk1, cx->strep, cx->M, cx->N);
    if (es->n_alpha + nadd > limit) return 1;
    if (orc_add_alpha_beta(toadd, nadd, es, pk, cx)) return -1;
    for (int a = 0; a < cx->N; a++)
    {
      if (act_old[a] < 2 && es->active[a] >= 2) return 1;
      if (act_old[a] >= 2 && es->active[a] > act_old[a]) return 1;
    }
    orc_update_bend_pts(es, pk1, cx);
    memcpy(act_old, es->active, sizeof(int) * NA);
    pk[0] = pk1[0];
    pk[1] = pk1[1];
  }
  return 0;
}

/* KinodynamicSearch::entangleCheckGivenPwp (kinodynamic_search.cpp:897-985). Faithful to the
 * reference quirk: the function returns inside the interval loop, so only interval 0 is checked
 * (:899, :982-983).  es is updated in place (ent_state_begin is passed by reference there). */
int orc_entangle_check_pwp(orc_ent* es, const orc_ectx* cx, int n, const double* cxy, const double* samp,
                           const unsigned char* known, int num_pol, int S, double T)
{
  if (n <= 0) return 0;
  const int NA = cx->N + cx->M, tcap = 4 * NA + 16;
  int* toadd = (int*)malloc(sizeof(int) * 2 * tcap);
  int* act_old = (int*)malloc(sizeof(int) * NA);
  int r = ent_interval_pass(es, cx, n, cxy, 0, samp, known, num_pol, S, T, 3 * NA, toadd, tcap, act_old);
  free(toadd);
  free(act_old);
  return r;
}

/* Chain of KinodynamicSearch::entanglesWithOtherAgents (kinodynamic_search.cpp:707-895) along a
 * given path: the producer of entStateVec (recoverEntStateVector :582-603).  List bound N+M
 * (:844-848); the tether-length test (:884-891) is not part of this chain.  Returns the number of
 * intervals processed before an entangling step (n if none), or -1 on capacity overflow. */
int orc_entangle_rollout(const orc_ent* es0, const orc_ectx* cx, int n, const double* cxy,
                         const double* samp, const unsigned char* known, int num_pol, int S, double T,
                         int* out_cnt, int* out_alpha, double* out_beta, int* out_bend, int* out_active)
{
  const int NA = cx->N + cx->M, cap = cx->cap, tcap = 4 * NA + 16;
  orc_ent es;
  es.alpha = (int*)malloc(sizeof(int) * 2 * cap);
  es.beta = (double*)malloc(sizeof(double) * cap);
  es.bend = (int*)malloc(sizeof(int) * cap);
  es.active = (int*)malloc(sizeof(int) * NA);
  es.n_alpha = es0->n_alpha;
  es.n_bend = es0->n_bend;
  memcpy(es.alpha, es0->alpha, sizeof(int) * 2 * cap);
  memcpy(es.beta, es0->beta, sizeof(double) * cap);
  memcpy(es.bend, es0->bend, sizeof(int) * cap);
  memcpy(es.active, es0->active, sizeof(int) * NA);
  int* toadd = (int*)malloc(sizeof(int) * 2 * tcap);
  int* act_old = (int*)malloc(sizeof(int) * NA);
  int done = n;
  for (int i = 0; i <= n; i++)
  {
    if (i > 0)
    {
      int r = ent_interval_pass(&es, cx, n, cxy, i - 1, samp, known, num_pol, S, T, NA, toadd, tcap, act_old);
      if (r < 0)
      {
        done = -1;
        break;
      }
      if (r > 0 && done == n) done = i - 1;
    }
    out_cnt[2 * i] = es.n_alpha;
    out_cnt[2 * i + 1] = es.n_bend;
    memcpy(out_alpha + (size_t)i * 2 * cap, es.alpha, sizeof(int) * 2 * cap);
    memcpy(out_beta + (size_t)i * cap, es.beta, sizeof(double) * cap);
    memcpy(out_bend + (size_t)i * cap, es.bend, sizeof(int) * cap);
    memcpy(out_active + (size_t)i * NA, es.active, sizeof(int) * NA);
  }
  free(es.alpha);
  free(es.beta);
  free(es.bend);
  free(es.active);
  free(toadd);
  free(act_old);
  return done;
}

/* ------------------------------------------------------------------------- */
/* back end: model build + interior point solve                               */
/* ------------------------------------------------------------------------- */

typedef struct
{
  int blk;      /* interval the row lives in */
  double c[12]; /* coefficients on (x[4], y[4], z[4]) of that interval */
  double rhs;   /* c . x_blk <= rhs */
} qrow;

typedef struct
{
  int n, nv, me, m; /* intervals, variables 12n, equalities, linear inequality rows */
  double T, W;
  double qp[4], qv[4], qa[4];
  double pf[3];
  int fallback, has_qc;
  double* Aeq; /* me x nv dense */
  double* beq;
  qrow* rows;
  double Pdiag_a; /* 72 T on every 'a' coefficient */
} qmodel;

static inline int vidx(int i, int ax, int r) { return i * 12 + ax * 4 + r; }

static inline double row_value(const qrow* r, const double* x)
{
  const double* xb = x + r->blk * 12;
  double v = 0;
  for (int k = 0; k < 12; k++) v += r->c[k] * xb[k];
  return v - r->rhs;
}

/* PolySolverGurobi::addObjective (:322-383) + addConstraints (:385-471, :659-708) given the
 * accepted separating lines.  Line rows follow :485-489 / :546-550 / :587-591 / :754-758. */
static void build_model(qmodel* md, const orc_params* par, const orc_replan_in* in, int fallback,
                        const double* lines, const unsigned char* line_ok, int LS, const double Ainv[16],
                        const double V[9])
{
  const int n = in->n;
  const double T = par->T_span;
  md->n = n;
  md->nv = 12 * n;
  md->T = T;
  md->W = par->weight;
  md->fallback = fallback;
  md->Pdiag_a = 72.0 * T;
  md->qp[0] = T * T * T, md->qp[1] = T * T, md->qp[2] = T, md->qp[3] = 1.0;         /* :126 */
  md->qv[0] = 3 * T * T, md->qv[1] = 2 * T, md->qv[2] = 1.0, md->qv[3] = 0.0;       /* :128 */
  md->qa[0] = 6 * T, md->qa[1] = 2.0, md->qa[2] = 0.0, md->qa[3] = 0.0;             /* :129 */
  const double* ci = in->coeff_init;
  for (int ax = 0; ax < 3; ax++) /* final_pos_ :226-228 */
  {
    const double* c = ci + ax * 32 + 4 * (n - 1);
    md->pf[ax] = md->qp[0] * c[0] + md->qp[1] * c[1] + md->qp[2] * c[2] + md->qp[3] * c[3];
  }
  {
    double dx = ci[3] - md->pf[0], dy = ci[32 + 3] - md->pf[1], dz = ci[64 + 3] - md->pf[2];
    md->has_qc = sqrt(dx * dx + dy * dy + dz * dz) < 1.0; /* :697-702 */
  }
  /* equalities */
  md->me = 9 + 9 * (n - 1) + (fallback ? 0 : 6);
  md->Aeq = (double*)calloc((size_t)md->me * md->nv, sizeof(double));
  md->beq = (double*)calloc(md->me, sizeof(double));
  int e = 0;
  for (int k1 = 1; k1 < 4; k1++) /* :390-396 */
    for (int ax = 0; ax < 3; ax++)
    {
      md->Aeq[e * md->nv + vidx(0, ax, k1)] = 1.0;
      md->beq[e] = ci[ax * 32 + k1];
      e++;
    }
  for (int i = 0; i < n - 1; i++) /* :400-425 */
    for (int ax = 0; ax < 3; ax++)
    {
      for (int k = 0; k < 4; k++) md->Aeq[e * md->nv + vidx(i, ax, k)] = md->qp[k];
      md->Aeq[e * md->nv + vidx(i + 1, ax, 3)] = -1.0;
      e++;
      for (int k = 0; k < 3; k++) md->Aeq[e * md->nv + vidx(i, ax, k)] = md->qv[k];
      md->Aeq[e * md->nv + vidx(i + 1, ax, 2)] = -1.0;
      e++;
      for (int k = 0; k < 2; k++) md->Aeq[e * md->nv + vidx(i, ax, k)] = md->qa[k];
      md->Aeq[e * md->nv + vidx(i + 1, ax, 1)] = -2.0;
      e++;
    }
  if (!fallback) /* :660-678 */
    for (int ax = 0; ax < 3; ax++)
    {
      for (int k = 0; k < 3; k++) md->Aeq[e * md->nv + vidx(n - 1, ax, k)] = md->qv[k];
      e++;
      for (int k = 0; k < 2; k++) md->Aeq[e * md->nv + vidx(n - 1, ax, k)] = md->qa[k];
      e++;
    }
  /* inequalities */
  int nl = 0;
  for (int i = 0; i < n; i++)
    for (int s = 0; s < LS; s++)
      if (line_ok[i * LS + s] == 1) nl++;
  md->rows = (qrow*)calloc((size_t)48 * n + 4 * nl + 1, sizeof(qrow));
  int m = 0;
  for (int i = 0; i < n; i++)
  {
    for (int ax = 0; ax < 3; ax++) /* :437-471 */
    {
      for (int k = 0; k < 4; k++)
      {
        qrow* r = &md->rows[m++];
        r->blk = i;
        for (int q = 0; q < 4; q++) r->c[ax * 4 + q] = Ainv[q * 4 + k];
        r->rhs = par->lim_max[ax];
        qrow* r2 = &md->rows[m++];
        r2->blk = i;
        for (int q = 0; q < 4; q++) r2->c[ax * 4 + q] = -Ainv[q * 4 + k];
        r2->rhs = -par->lim_min[ax];
      }
      for (int k = 0; k < 3; k++)
      {
        qrow* r = &md->rows[m++];
        r->blk = i;
        for (int q = 0; q < 3; q++) r->c[ax * 4 + q] = V[q * 3 + k];
        r->rhs = par->v_max;
        qrow* r2 = &md->rows[m++];
        r2->blk = i;
        for (int q = 0; q < 3; q++) r2->c[ax * 4 + q] = -V[q * 3 + k];
        r2->rhs = par->v_max;
      }
      qrow* r = &md->rows[m++];
      r->blk = i;
      r->c[ax * 4 + 0] = 6.0 * T;
      r->c[ax * 4 + 1] = 2.0;
      r->rhs = par->a_max;
      qrow* r2 = &md->rows[m++];
      r2->blk = i;
      r2->c[ax * 4 + 0] = -6.0 * T;
      r2->c[ax * 4 + 1] = -2.0;
      r2->rhs = par->a_max;
    }
    for (int s = 0; s < LS; s++)
      if (line_ok[i * LS + s] == 1)
      {
        const double* l = lines + (size_t)(i * LS + s) * 3;
        for (int k = 0; k < 4; k++)
        {
          qrow* r = &md->rows[m++];
          r->blk = i;
          for (int q = 0; q < 4; q++)
          {
            r->c[q] = l[0] * Ainv[q * 4 + k];
            r->c[4 + q] = l[1] * Ainv[q * 4 + k];
          }
          r->rhs = 1.0 - l[2];
        }
      }
  }
  md->m = m;
}

static void free_model(qmodel* md)
{
  free(md->Aeq);
  free(md->beq);
  free(md->rows);
}

/* objective value: 36 T sum a^2 + W sum (qp.x - pf)^2 [+ W sum ((qv.x)^2 + (qa.x)^2)]  (:322-380) */
static double model_objective(const qmodel* md, const double* x)
{
  double f = 0.0;
  for (int i = 0; i < md->n; i++)
    for (int ax = 0; ax < 3; ax++)
    {
      double a = x[vidx(i, ax, 0)];
      f += 36.0 * md->T * a * a;
    }
  for (int ax = 0; ax < 3; ax++)
  {
    const double* c = x + vidx(md->n - 1, ax, 0);
    double e = md->qp[0] * c[0] + md->qp[1] * c[1] + md->qp[2] * c[2] + md->qp[3] * c[3] - md->pf[ax];
    f += md->W * e * e;
    if (md->fallback)
    {
      double v = md->qv[0] * c[0] + md->qv[1] * c[1] + md->qv[2] * c[2];
      double a = md->qa[0] * c[0] + md->qa[1] * c[1];
      f += md->W * (v * v + a * a);
    }
  }
  return f;
}

/* gradient of the objective g = P x + q, and P block of the last interval added into K */
static void model_grad(const qmodel* md, const double* x, double* g)
{
  for (int v = 0; v < md->nv; v++) g[v] = 0.0;
  for (int i = 0; i < md->n; i++)
    for (int ax = 0; ax < 3; ax++) g[vidx(i, ax, 0)] += md->Pdiag_a * x[vidx(i, ax, 0)];
  for (int ax = 0; ax < 3; ax++)
  {
    const int o = vidx(md->n - 1, ax, 0);
    const double* c = x + o;
    double e = md->qp[0] * c[0] + md->qp[1] * c[1] + md->qp[2] * c[2] + md->qp[3] * c[3] - md->pf[ax];
    for (int k = 0; k < 4; k++) g[o + k] += 2.0 * md->W * e * md->qp[k];
    if (md->fallback)
    {
      double v = md->qv[0] * c[0] + md->qv[1] * c[1] + md->qv[2] * c[2];
      double a = md->qa[0] * c[0] + md->qa[1] * c[1];
      for (int k = 0; k < 3; k++) g[o + k] += 2.0 * md->W * v * md->qv[k];
      for (int k = 0; k < 2; k++) g[o + k] += 2.0 * md->W * a * md->qa[k];
    }
  }
}

static void model_add_P(const qmodel* md, double* K /* n blocks of 12x12 */)
{
  for (int i = 0; i < md->n; i++)
    for (int ax = 0; ax < 3; ax++) K[i * 144 + (ax * 4) * 12 + ax * 4] += md->Pdiag_a;
  double* Kl = K + (md->n - 1) * 144;
  for (int ax = 0; ax < 3; ax++)
    for (int a = 0; a < 4; a++)
      for (int b = 0; b < 4; b++)
      {
        double v = 2.0 * md->W * md->qp[a] * md->qp[b];
        if (md->fallback) v += 2.0 * md->W * (md->qv[a] * md->qv[b] + md->qa[a] * md->qa[b]);
        Kl[(ax * 4 + a) * 12 + ax * 4 + b] += v;
      }
}

/* terminal quadratic constraint value and gradient (last block only): :681-702 */
static double qc_value(const qmodel* md, const double* x, double e3[3])
{
  double c = -0.10 * 0.10;
  for (int ax = 0; ax < 3; ax++)
  {
    const double* q = x + vidx(md->n - 1, ax, 0);
    e3[ax] = md->qp[0] * q[0] + md->qp[1] * q[1] + md->qp[2] * q[2] + md->qp[3] * q[3] - md->pf[ax];
    c += e3[ax] * e3[ax];
  }
  return c;
}

/*
 * Null space of the equality rows by Householder QR of A^T with column pivoting (what a generic
 * solver's presolve does with dense equalities).  On return: Z (nv x nz, orthonormal columns),
 * xp = minimum-norm solution of A x = b.  Returns 0, or 1 when the equalities are inconsistent.
 */
static int eq_nullspace(const double* A, const double* b, int me, int nv, double* Z, double* xp, int* nz_out)
{
  double* M = (double*)malloc(sizeof(double) * nv * me); /* M = A^T, nv x me, column j = row j of A */
  double* vs = (double*)calloc((size_t)nv * (me + 1), sizeof(double));
  double* beta = (double*)calloc(me + 1, sizeof(double));
  int* perm = (int*)malloc(sizeof(int) * me);
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < me; j++) M[i * me + j] = A[j * nv + i];
  for (int j = 0; j < me; j++) perm[j] = j;
  double norm0 = 0.0;
  int rank = 0;
  const int kmax = nv < me ? nv : me;
  for (int k = 0; k < kmax; k++)
  {
    int best = k;
    double bn = -1.0;
    for (int j = k; j < me; j++)
    {
      double s = 0;
      for (int i = k; i < nv; i++) s += M[i * me + j] * M[i * me + j];
      if (s > bn)
      {
        bn = s;
        best = j;
      }
    }
    bn = sqrt(bn);
    if (k == 0) norm0 = bn;
    if (!(bn > 1e-11 * (norm0 > 0 ? norm0 : 1.0))) break;
    if (best != k)
    {
      for (int i = 0; i < nv; i++)
      {
        double t = M[i * me + k];
        M[i * me + k] = M[i * me + best];
        M[i * me + best] = t;
      }
      int t = perm[k];
      perm[k] = perm[best];
      perm[best] = t;
    }
    /* Householder vector for column k */
    double alpha = M[k * me + k] >= 0 ? -bn : bn;
    double* v = vs + (size_t)k * nv;
    for (int i = 0; i < nv; i++) v[i] = (i < k) ? 0.0 : M[i * me + k];
    v[k] -= alpha;
    double vn = 0;
    for (int i = k; i < nv; i++) vn += v[i] * v[i];
    beta[k] = vn > 0 ? 2.0 / vn : 0.0;
    for (int j = k; j < me; j++)
    {
      double d = 0;
      for (int i = k; i < nv; i++) d += v[i] * M[i * me + j];
      d *= beta[k];
      for (int i = k; i < nv; i++) M[i * me + j] -= d * v[i];
    }
    rank = k + 1;
  }
  /* u[:rank] from R1^T u = (Pi^T b)[:rank] (forward substitution), u[rank:] = 0 ; x = Q u */
  double* u = (double*)calloc(nv, sizeof(double));
  for (int i = 0; i < rank; i++)
  {
    double s = b[perm[i]];
    for (int k = 0; k < i; k++) s -= M[k * me + i] * u[k];
    u[i] = s / M[i * me + i];
  }
  for (int k = rank - 1; k >= 0; k--)
  {
    const double* v = vs + (size_t)k * nv;
    double d = 0;
    for (int i = k; i < nv; i++) d += v[i] * u[i];
    d *= beta[k];
    for (int i = k; i < nv; i++) u[i] -= d * v[i];
  }
  memcpy(xp, u, sizeof(double) * nv);
  const int nz = nv - rank;
  for (int c = 0; c < nz; c++)
  {
    for (int i = 0; i < nv; i++) u[i] = (i == rank + c) ? 1.0 : 0.0;
    for (int k = rank - 1; k >= 0; k--)
    {
      const double* v = vs + (size_t)k * nv;
      double d = 0;
      for (int i = k; i < nv; i++) d += v[i] * u[i];
      d *= beta[k];
      for (int i = k; i < nv; i++) u[i] -= d * v[i];
    }
    for (int i = 0; i < nv; i++) Z[i * nz + c] = u[i];
  }
  *nz_out = nz;
  int bad = 0;
  double bmax = 0;
  for (int e = 0; e < me; e++)
    if (fabs(b[e]) > bmax) bmax = fabs(b[e]);
  for (int e = 0; e < me; e++)
  {
    double r = -b[e];
    for (int i = 0; i < nv; i++) r += A[e * nv + i] * xp[i];
    if (fabs(r) > 1e-9 * (1.0 + bmax)) bad = 1;
  }
  free(M);
  free(vs);
  free(beta);
  free(perm);
  free(u);
  return bad;
}

/*
 * Stand-in for m_.optimize() (solver_gurobi_poly.cpp:823, :846).  Gurobi is not available; the
 * model is a convex QP (QCQP when :699 fires) whose Hessian is positive definite on the null space
 * of the equalities, so the minimiser is unique.  Solved the way a generic solver would: the dense
 * equality rows are eliminated numerically (eq_nullspace), then a textbook infeasible-start
 * Mehrotra predictor-corrector interior-point method (Nocedal & Wright, Alg. 16.4) runs on the
 * remaining dense inequality rows.  Returns 1 when converged within max_iter (== "a solution
 * exists", :832-836), else 0 (infeasible / no solution).
 */
static int ipm_solve(const qmodel* md, double* x, int max_iter, double tol, int* iters_out)
{
  const int n = md->n, nv = md->nv, me = md->me, m = md->m, mq = m + (md->has_qc ? 1 : 0);
  double* Z = (double*)malloc(sizeof(double) * nv * nv);
  double* xp = (double*)malloc(sizeof(double) * nv);
  int nz = 0, it = 0, converged = 0;
  double mu_div = 1e300;
  if (iters_out) *iters_out = 0;
  if (eq_nullspace(md->Aeq, md->beq, me, nv, Z, xp, &nz))
  {
    free(Z);
    free(xp);
    return 0;
  }
  /* reduced rows Gr = G Z, hr = h - G xp */
  double* Gr = (double*)malloc(sizeof(double) * (size_t)(m + 1) * (nz + 1));
  double* hr = (double*)malloc(sizeof(double) * (m + 1));
  double hn = 0.0;
  for (int r = 0; r < m; r++)
  {
    const qrow* q = &md->rows[r];
    double v = q->rhs;
    for (int k = 0; k < 12; k++) v -= q->c[k] * xp[q->blk * 12 + k];
    hr[r] = v;
    if (fabs(q->rhs) > hn) hn = fabs(q->rhs);
    for (int c = 0; c < nz; c++)
    {
      double a = 0;
      for (int k = 0; k < 12; k++) a += q->c[k] * Z[(q->blk * 12 + k) * nz + c];
      Gr[(size_t)r * nz + c] = a;
    }
  }
  if (nz == 0)
  { /* the equalities leave a single point */
    int ok = 1;
    for (int r = 0; r < m; r++)
      if (hr[r] < -1e-9 * (1.0 + hn)) ok = 0;
    double e3[3];
    if (md->has_qc && qc_value(md, xp, e3) > 1e-9) ok = 0;
    if (ok) memcpy(x, xp, sizeof(double) * nv);
    free(Z), free(xp), free(Gr), free(hr);
    return ok;
  }
  /* reduced Hessian Hr = Z^T P Z (constant) */
  double* Pfull = (double*)calloc((size_t)n * 144, sizeof(double));
  model_add_P(md, Pfull);
  double* Hr = (double*)calloc((size_t)nz * nz, sizeof(double));
  double* PZ = (double*)calloc((size_t)nv * nz, sizeof(double));
  for (int b = 0; b < n; b++)
    for (int a = 0; a < 12; a++)
      for (int k = 0; k < 12; k++)
      {
        double p = Pfull[b * 144 + a * 12 + k];
        if (p != 0.0)
          for (int c = 0; c < nz; c++) PZ[(b * 12 + a) * nz + c] += p * Z[(b * 12 + k) * nz + c];
      }
  for (int i = 0; i < nv; i++)
    for (int a = 0; a < nz; a++)
    {
      double za = Z[i * nz + a];
      if (za != 0.0)
        for (int c = 0; c < nz; c++) Hr[a * nz + c] += za * PZ[i * nz + c];
    }
  /* terminal-position row of the quadratic constraint in reduced coordinates: tq[ax] . w */
  double* tq = (double*)calloc((size_t)3 * nz, sizeof(double));
  for (int ax = 0; ax < 3; ax++)
    for (int k = 0; k < 4; k++)
      for (int c = 0; c < nz; c++) tq[ax * nz + c] += md->qp[k] * Z[(vidx(n - 1, ax, k)) * nz + c];

  double* w = (double*)calloc(nz, sizeof(double));
  double* buf = (double*)malloc(sizeof(double) * (8 * (size_t)mq + 8 * (size_t)nz + (size_t)nz * nz + nv));
  double *s = buf, *lam = s + mq, *rp = lam + mq, *ds = rp + mq, *dl = ds + mq, *dsa = dl + mq, *dla = dsa + mq,
         *rc = dla + mq;
  double *g = rc + mq, *rd = g + nz, *rhs = rd + nz, *dw = rhs + nz, *gq = dw + nz;
  double* K = buf + 8 * (size_t)mq + 8 * (size_t)nz;
  double* xcur = (double*)malloc(sizeof(double) * nv);
  double* gx = (double*)malloc(sizeof(double) * nv);

  /* start: w = Z^T (x_frontend - xp) (projection of the front-end path), s = max(h - G w, 1), lam = 1 */
  for (int c = 0; c < nz; c++)
  {
    double a = 0;
    for (int i = 0; i < nv; i++) a += Z[i * nz + c] * (x[i] - xp[i]);
    w[c] = a;
  }
#define XCUR()                                                      \
  for (int i = 0; i < nv; i++)                                      \
  {                                                                 \
    double a = xp[i];                                               \
    for (int c = 0; c < nz; c++) a += Z[i * nz + c] * w[c];         \
    xcur[i] = a;                                                    \
  }
  XCUR();
  for (int r = 0; r < m; r++)
  {
    double v = hr[r];
    for (int c = 0; c < nz; c++) v -= Gr[(size_t)r * nz + c] * w[c];
    s[r] = v > 1.0 ? v : 1.0;
    lam[r] = 1.0;
  }
  if (md->has_qc)
  {
    double e3[3];
    double v = -qc_value(md, xcur, e3);
    s[m] = v > 1.0 ? v : 1.0;
    lam[m] = 1.0;
  }

  for (it = 0; it <= max_iter; it++)
  {
    const int init_pass = (it == 0);
    double e3[3] = { 0, 0, 0 };
    XCUR();
    model_grad(md, xcur, gx);
    for (int c = 0; c < nz; c++)
    {
      double a = 0;
      for (int i = 0; i < nv; i++) a += Z[i * nz + c] * gx[i];
      g[c] = a;
      rd[c] = a;
    }
    for (int r = 0; r < m; r++)
    {
      double v = -hr[r];
      const double* gr = Gr + (size_t)r * nz;
      for (int c = 0; c < nz; c++)
      {
        v += gr[c] * w[c];
        rd[c] += gr[c] * lam[r];
      }
      rp[r] = v + s[r];
    }
    for (int c = 0; c < nz; c++) gq[c] = 0.0;
    if (md->has_qc)
    {
      rp[m] = qc_value(md, xcur, e3) + s[m];
      for (int ax = 0; ax < 3; ax++)
        for (int c = 0; c < nz; c++) gq[c] += 2.0 * e3[ax] * tq[ax * nz + c];
      for (int c = 0; c < nz; c++) rd[c] += gq[c] * lam[m];
    }
    double mu = 0.0, rpn = 0.0, rdn = 0.0, gn = 0.0;
    for (int r = 0; r < mq; r++)
    {
      mu += s[r] * lam[r];
      if (fabs(rp[r]) > rpn) rpn = fabs(rp[r]);
    }
    mu /= mq;
    for (int c = 0; c < nz; c++)
    {
      if (fabs(rd[c]) > rdn) rdn = fabs(rd[c]);
      if (fabs(g[c]) > gn) gn = fabs(g[c]);
    }
    const double fobj = model_objective(md, xcur);
    if (getenv("ORC_DEBUG")) fprintf(stderr, "it %d mu %.3e rp %.3e rd %.3e f %.12g\n", it, mu, rpn, rdn, fobj);
    if (!init_pass)
    {
      /* scaled stopping test: every residual relative to the size of the quantities it is made of */
      if (rpn <= tol * (1.0 + hn) && rdn <= tol * (1.0 + gn) && mu * mq <= tol * (1.0 + fabs(fobj)))
      {
        converged = 1;
        break;
      }
      if (it == max_iter) break;
      if (!(mu == mu) || !(rpn == rpn) || !(rdn == rdn)) break; /* NaN */
      /* diverged (infeasible model: the multipliers run away long before anything overflows): same outcome as the NaN
         test and the iteration cap, sooner; a converging solve never leaves mu 1e12 above where it started */
      if (it == 1) mu_div = 1e12 * (1.0 + mu);
      if (mu > mu_div) break;
    }
    /* K = Hr + lam_q Hess(c) + Gr^T D Gr */
    memcpy(K, Hr, sizeof(double) * nz * nz);
    for (int r = 0; r < m; r++)
    {
      const double* gr = Gr + (size_t)r * nz;
      double d = lam[r] / s[r];
      for (int a = 0; a < nz; a++)
      {
        double da = d * gr[a];
        if (da == 0.0) continue;
        for (int c = 0; c <= a; c++) K[a * nz + c] += da * gr[c];
      }
    }
    if (md->has_qc)
    {
      double d = lam[m] / s[m];
      for (int a = 0; a < nz; a++)
        for (int c = 0; c <= a; c++)
        {
          double hq = 0;
          for (int ax = 0; ax < 3; ax++) hq += 2.0 * tq[ax * nz + a] * tq[ax * nz + c];
          K[a * nz + c] += d * gq[a] * gq[c] + lam[m] * hq;
        }
    }
    for (int a = 0; a < nz; a++)
      for (int c = a + 1; c < nz; c++) K[a * nz + c] = K[c * nz + a];
    chol_factor(K, nz, nz);

    double sigma = 0.0;
    for (int pass = 0; pass < 2; pass++)
    {
      for (int r = 0; r < mq; r++)
        rc[r] = (pass == 0) ? s[r] * lam[r] : s[r] * lam[r] + dsa[r] * dla[r] - sigma * mu;
      for (int c = 0; c < nz; c++) rhs[c] = -rd[c];
      for (int r = 0; r < m; r++)
      {
        const double* gr = Gr + (size_t)r * nz;
        double wv = (lam[r] * rp[r] - rc[r]) / s[r];
        for (int c = 0; c < nz; c++) rhs[c] -= gr[c] * wv;
      }
      if (md->has_qc)
      {
        double wv = (lam[m] * rp[m] - rc[m]) / s[m];
        for (int c = 0; c < nz; c++) rhs[c] -= gq[c] * wv;
      }
      memcpy(dw, rhs, sizeof(double) * nz);
      chol_solve(K, nz, nz, dw);
      for (int r = 0; r < m; r++)
      {
        const double* gr = Gr + (size_t)r * nz;
        double gd = 0;
        for (int c = 0; c < nz; c++) gd += gr[c] * dw[c];
        ds[r] = -rp[r] - gd;
      }
      if (md->has_qc)
      {
        double gd = 0;
        for (int c = 0; c < nz; c++) gd += gq[c] * dw[c];
        ds[m] = -rp[m] - gd;
      }
      for (int r = 0; r < mq; r++) dl[r] = (-rc[r] - lam[r] * ds[r]) / s[r];
      if (init_pass)
      { /* Nocedal-Wright starting-point heuristic: s0 = max(1,|s+ds_aff|), lam0 likewise */
        for (int r = 0; r < mq; r++)
        {
          double a = fabs(s[r] + ds[r]), b2 = fabs(lam[r] + dl[r]);
          s[r] = a > 1.0 ? a : 1.0;
          lam[r] = b2 > 1.0 ? b2 : 1.0;
        }
        break;
      }
      double amax = 1e300; /* largest step keeping (s, lam) >= 0 */
      for (int r = 0; r < mq; r++)
      {
        if (ds[r] < 0)
        {
          double a = -s[r] / ds[r];
          if (a < amax) amax = a;
        }
        if (dl[r] < 0)
        {
          double a = -lam[r] / dl[r];
          if (a < amax) amax = a;
        }
      }
      if (pass == 0)
      {
        double alpha = amax < 1.0 ? amax : 1.0;
        double mua = 0;
        for (int r = 0; r < mq; r++) mua += (s[r] + alpha * ds[r]) * (lam[r] + alpha * dl[r]);
        mua /= mq;
        sigma = (mua / mu) * (mua / mu) * (mua / mu);
        /* do not drive the complementarity below a tenth of what the stopping test needs: at mu ~ 1e-12
         * the normal matrix is too ill-conditioned for the dual residual to reach its tolerance */
        {
          const double mu_floor = 0.1 * tol * (1.0 + fabs(fobj)) / mq;
          if (sigma * mu < mu_floor) sigma = mu_floor / mu;
        }
        memcpy(dsa, ds, sizeof(double) * mq);
        memcpy(dla, dl, sizeof(double) * mq);
      }
      else
      {
        double eta = 1.0 - 1.0 / ((it + 3.0) * (it + 3.0)); /* fraction to the boundary -> 1 */
        double a = eta * amax;
        if (a > 1.0) a = 1.0;
        for (int c = 0; c < nz; c++) w[c] += a * dw[c];
        for (int r = 0; r < mq; r++)
        {
          s[r] += a * ds[r];
          lam[r] += a * dl[r];
        }
      }
    }
  }
  if (converged)
  {
    XCUR();
    memcpy(x, xcur, sizeof(double) * nv);
  }
#undef XCUR
  if (iters_out) *iters_out = it;
  free(Z), free(xp), free(Gr), free(hr), free(Pfull), free(Hr), free(PZ), free(tq), free(w), free(buf);
  free(xcur), free(gx);
  return converged;
}

/* ---- LP generation: the separation part of addConstraints ---- */

static double dist2(const double* a, const double* b)
{
  double dx = a[0] - b[0], dy = a[1] - b[1];
  return sqrt(dx * dx + dy * dy);
}

/* control points of the initial path: ctrlPtsInit_ (:232-243), out[i][k][2] */
static void ctrl_points_init(const double* ci, int n, const double Ainv[16], double* out)
{
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 4; k++)
    {
      double x = 0, y = 0;
      for (int r = 0; r < 4; r++)
      {
        x += ci[4 * i + r] * Ainv[r * 4 + k];
        y += ci[32 + 4 * i + r] * Ainv[r * 4 + k];
      }
      out[(i * 4 + k) * 2] = x;
      out[(i * 4 + k) * 2 + 1] = y;
    }
}

/* returns 0, or -3 when more entangle LPs are generated than ent_slots */
static int generate_lines(const orc_params* par, const orc_replan_in* in, const double Ainv[16], double* lines,
                          unsigned char* line_ok, int LS)
{
  const int n = in->n, N = par->num_agents, M = par->num_static, NH = in->n_hull_slots;
  const int cap = par->ent_cap, NA = N + M;
  double cp[ORC_NPOL_MAX * 4 * 2];
  ctrl_points_init(in->coeff_init, n, Ainv, cp);
  memset(line_ok, 0, (size_t)ORC_NPOL_MAX * LS);
  const double long_len = sqrt((par->lim_max[0] - par->lim_min[0]) * (par->lim_max[0] - par->lim_min[0]) +
                               (par->lim_max[1] - par->lim_min[1]) * (par->lim_max[1] - par->lim_min[1])); /* :173 */
  for (int i = 0; i < n; i++)
  {
    const double* Bp = cp + i * 8;
    double* L = lines + (size_t)i * LS * 3;
    unsigned char* ok = line_ok + (size_t)i * LS;
    /* other agents :477-495 */
    for (int s = 0; s < NH; s++)
    {
      long long o0 = in->hull_ptr[s * 8 + i], o1 = in->hull_ptr[s * 8 + i + 1];
      int cnt = (int)(o1 - o0);
      if (cnt <= 0) continue;
      ok[s] = orc_separate(in->hull_xy + 2 * o0, cnt, Bp, 4, L + 3 * s) ? 1 : 2;
    }
    /* bases :521-553 (includes the agent's own base) */
    const double base_radius = 0.7;
    for (int j = 0; j < N; j++)
    {
      int close = 0;
      for (int k = 0; k < 4; k++)
        if (dist2(Bp + 2 * k, in->pb + 2 * j) < base_radius * 3)
        {
          close = 1;
          break;
        }
      if (!close) continue;
      const double bx = in->pb[2 * j], by = in->pb[2 * j + 1];
      const double hull[8] = { bx + base_radius, by + base_radius, bx + base_radius, by - base_radius,
                               bx - base_radius, by + base_radius, bx - base_radius, by - base_radius };
      ok[NH + j] = orc_separate(hull, 4, Bp, 4, L + 3 * (NH + j)) ? 1 : 2;
    }
    /* static obstacles :556-593 */
    for (int j = 0; j < M; j++)
    {
      const double* sv = in->st_xy + 2 * in->st_ptr[j];
      int cnt = (int)(in->st_ptr[j + 1] - in->st_ptr[j]);
      int close = 0;
      double dist = dist2(Bp, sv);
      for (int k = 0; k < 3; k++)
      {
        dist -= dist2(Bp + 2 * (k + 1), Bp + 2 * k);
        if (dist < 0)
        {
          close = 1;
          break;
        }
      }
      for (int k = 0; k < cnt - 1; k++)
      {
        dist -= dist2(sv + 2 * (k + 1), sv + 2 * k);
        if (dist < 0)
        {
          close = 1;
          break;
        }
      }
      if (!close) continue;
      ok[NH + N + j] = orc_separate(sv, cnt, Bp, 4, L + 3 * (NH + N + j)) ? 1 : 2;
    }
    /* non-entangling :620-642 -> addEntangleConstraintForIJCase :715-784 */
    int eslot = 0;
    const int* alpha = in->esv_alpha + (size_t)i * cap * 2;
    const int n_alpha = in->esv_cnt[2 * i];
    const int* active = in->esv_active + (size_t)i * NA;
    double hulldist = 0.0;
    for (int k = 0; k < 3; k++) hulldist += dist2(Bp + 2 * (k + 1), Bp + 2 * k); /* :738-742 */
    for (int j = 0; j < N; j++)
    {
      if (j == in->agent_id - 1) continue;
      if (active[j] != 1) continue;
      int case_id = 0;
      for (int jj = 0; jj < n_alpha; jj++)
        if (alpha[2 * jj] == j + 1) case_id = alpha[2 * jj + 1];
      if (case_id == 0) continue;
      const int nb = in->bp_cnt[j];
      const double* bend = in->bp_xy + (size_t)2 * par->bp_max * j;
      const double* posj = in->nih0 + ((size_t)j * 8 + i) * 2; /* hullsNoInflation_[j][i].col(0) */
      for (int k = 1; k < nb + 1; k++)
      {
        if (k == case_id) continue;
        double pA[2], pB[2];
        if (nb < 1 || posj[0] != posj[0]) continue; /* unknown agent: nothing to constrain */
        if (k == 1) /* :719-724 */
        {
          pA[0] = (1 - long_len) * bend[2 * (nb - 1)] + long_len * posj[0];
          pA[1] = (1 - long_len) * bend[2 * (nb - 1) + 1] + long_len * posj[1];
          pB[0] = posj[0];
          pB[1] = posj[1];
        }
        else if (k > 1 && k <= nb) /* :725-730 */
        {
          pA[0] = bend[2 * (k - 2)];
          pA[1] = bend[2 * (k - 2) + 1];
          pB[0] = bend[2 * (k - 1)];
          pB[1] = bend[2 * (k - 1) + 1];
        }
        else
          continue;
        if (dist2(pA, Bp) - hulldist > 0 && dist2(pB, Bp) - hulldist > 0) continue; /* :743-745 */
        if (eslot >= par->ent_slots) return -3;
        const double Aset[4] = { pA[0], pA[1], pB[0], pB[1] };
        int sl = NH + N + M + eslot;
        ok[sl] = orc_separate(Aset, 2, Bp, 4, L + 3 * sl) ? 1 : 2; /* :751 (4-arg overload) */
        eslot++;
      }
    }
  }
  return 0;
}

/*
 * PolySolverGurobi::optimize (solver_gurobi_poly.cpp:804-887) for one agent, preceded by the
 * setters' bookkeeping (:187-320).  Status path: direct / fallback / failed (:832-861);
 * z override (:879-880); objective value (:882).
 */
int orc_replan(const orc_params* par, const orc_replan_in* in, orc_replan_out* out)
{
  const int n = in->n, N = par->num_agents, M = par->num_static;
  const int LS = in->n_hull_slots + N + M + par->ent_slots;
  double Ainv[16], V[9];
  orc_basis(par->T_span, Ainv, V, 0);
  double* lines = out->lines;
  unsigned char* line_ok = out->line_ok;
  int own = 0;
  if (!lines)
  {
    lines = (double*)malloc(sizeof(double) * ORC_NPOL_MAX * LS * 3);
    line_ok = (unsigned char*)malloc((size_t)ORC_NPOL_MAX * LS);
    own = 1;
  }
  int rc = generate_lines(par, in, Ainv, lines, line_ok, LS);
  int status = ORC_STATUS_FAILED;
  double x[12 * ORC_NPOL_MAX];
  out->iters[0] = out->iters[1] = 0;
  *out->obj = 0.0;
  if (rc == 0)
    for (int attempt = 0; attempt < 2; attempt++)
    {
      qmodel md;
      build_model(&md, par, in, attempt, lines, line_ok, LS, Ainv, V);
      for (int i = 0; i < n; i++)
        for (int ax = 0; ax < 3; ax++)
          for (int r = 0; r < 4; r++) x[vidx(i, ax, r)] = in->coeff_init[ax * 32 + 4 * i + r];
      /* The checker gets eight times the product's iteration cap: this full-space model keeps every redundant line, which
         slows the central path down on crowded scenes (40-128 iterations where the pruned model of the product needs 12-18,
         tests/test_crafted_branches.py::test_crowded_worlds_*); the cap only has to catch non-convergence, and a feasible
         model must not be declared unsolved by the checker because it is the slower of the two. */
      int ok = ipm_solve(&md, x, 8 * par->ipm_max_iter, par->ipm_tol, &out->iters[attempt]);
      if (ok)
      {
        *out->obj = model_objective(&md, x);
        status = attempt == 0 ? ORC_STATUS_OK : ORC_STATUS_FALLBACK;
      }
      free_model(&md);
      if (ok) break;
    }
  memcpy(out->coeff_out, in->coeff_init, sizeof(double) * 96); /* pwp_out_ = pwp_init_ :858 */
  if (status != ORC_STATUS_FAILED)
  {
    for (int i = 0; i < n; i++)
      for (int ax = 0; ax < 3; ax++)
        for (int r = 0; r < 4; r++) out->coeff_out[ax * 32 + 4 * i + r] = x[vidx(i, ax, r)];
    /* :879-880 */
    const double* ci = in->coeff_init;
    const double T = par->T_span;
    double pfx = 0, pfy = 0;
    const double qp[4] = { T * T * T, T * T, T, 1.0 };
    for (int r = 0; r < 4; r++)
    {
      pfx += qp[r] * ci[4 * (n - 1) + r];
      pfy += qp[r] * ci[32 + 4 * (n - 1) + r];
    }
    double dx = ci[3] - pfx, dy = ci[32 + 3] - pfy;
    if (sqrt(dx * dx + dy * dy) < 1.0) memcpy(out->coeff_out + 64, ci + 64, sizeof(double) * 32);
  }
  *out->status = status;
  if (own)
  {
    free(lines);
    free(line_ok);
  }
  return rc;
}

int orc_export_qp(const orc_params* par, const orc_replan_in* in, int fallback, const double* lines,
                  const unsigned char* line_ok, int LS, double* P, double* q, double* c0, double* Aeq,
                  double* beq, int* n_eq, double* G, double* h, int max_rows, int* has_qc)
{
  double Ainv[16], V[9];
  orc_basis(par->T_span, Ainv, V, 0);
  qmodel md;
  build_model(&md, par, in, fallback, lines, line_ok, LS, Ainv, V);
  const int nv = md.nv, n = md.n;
  if (md.m > max_rows)
  {
    free_model(&md);
    return -1;
  }
  memset(P, 0, sizeof(double) * nv * nv);
  double* K = (double*)calloc((size_t)n * 144, sizeof(double));
  model_add_P(&md, K);
  for (int b = 0; b < n; b++)
    for (int a = 0; a < 12; a++)
      for (int c = 0; c < 12; c++) P[(b * 12 + a) * nv + b * 12 + c] = K[b * 144 + a * 12 + c];
  free(K);
  double* zero = (double*)calloc(nv, sizeof(double));
  model_grad(&md, zero, q);
  *c0 = model_objective(&md, zero);
  free(zero);
  memcpy(Aeq, md.Aeq, sizeof(double) * md.me * nv);
  memcpy(beq, md.beq, sizeof(double) * md.me);
  *n_eq = md.me;
  memset(G, 0, sizeof(double) * (size_t)md.m * nv);
  for (int r = 0; r < md.m; r++)
  {
    for (int k = 0; k < 12; k++) G[(size_t)r * nv + md.rows[r].blk * 12 + k] = md.rows[r].c[k];
    h[r] = md.rows[r].rhs;
  }
  *has_qc = md.has_qc;
  int m = md.m;
  free_model(&md);
  return m;
}

/* PolySolverGurobi::generatePwpOut sampling loop (:911-934): states[k] = pos(3) vel(3) acc(3) jerk(3).
 * Operation order as written there: the power vectors tp = (dt^3, dt^2, dt, 1), tv = (3 dt^2, 2 dt, 1), ta = (6 dt, 2)
 * are formed first, then each state is a row of coefficients times that vector, summed from the left.
 * KinodynamicSearch::generatePwpOut (kinodynamic_search.cpp:621-668) is the same code; tests/test_reference_pin.py
 * compares this function with it bit for bit. */
int orc_generate_traj(const double* coeff, int n, double T, double dc, double* states, int max_states)
{
  double t = 0;
  int i = 0, cnt = 0;
  while (i < n && cnt < max_states)
  {
    const double dt = t - i * T;
    const double tp0 = dt * dt * dt, tp1 = dt * dt, tv0 = 3 * dt * dt, tv1 = 2 * dt, ta0 = 6 * dt;
    double* st = states + 12 * cnt;
    for (int ax = 0; ax < 3; ax++)
    {
      const double* c = coeff + ax * 32 + 4 * i;
      st[ax] = c[0] * tp0 + c[1] * tp1 + c[2] * dt + c[3];
      st[3 + ax] = c[0] * tv0 + c[1] * tv1 + c[2];
      st[6 + ax] = c[0] * ta0 + c[1] * 2;
      st[9 + ax] = c[0] * 6;
    }
    cnt++;
    t += dc;
    if (t > (i + 1) * T) i++;
  }
  return cnt;
}

int orc_replan_batch(const orc_params* par, const orc_batch* b, int nthreads)
{
  const int N = par->num_agents, M = par->num_static, NA = N + M, cap = par->ent_cap;
  const int NH = b->n_hull_slots, LS = NH + N + M + par->ent_slots;
  int rc_all = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int a = 0; a < b->B; a++)
  {
    orc_replan_in in;
    orc_replan_out out;
    in.agent_id = b->agent_id[a];
    in.n = b->n_int[a];
    in.coeff_init = b->coeff_init + (size_t)a * 96;
    in.n_hull_slots = NH;
    in.hull_ptr = b->hull_ptr + (size_t)a * NH * 8;
    in.hull_xy = b->hull_xy;
    in.nih0 = b->nih0 + (size_t)a * N * 16;
    in.st_ptr = b->st_ptr;
    in.st_xy = b->st_xy;
    in.esv_cnt = b->esv_cnt + (size_t)a * 18;
    in.esv_alpha = b->esv_alpha + (size_t)a * 9 * cap * 2;
    in.esv_active = b->esv_active + (size_t)a * 9 * NA;
    in.bp_cnt = b->bp_shared ? b->bp_cnt : b->bp_cnt + (size_t)a * N;
    in.bp_xy = b->bp_shared ? b->bp_xy : b->bp_xy + (size_t)a * N * par->bp_max * 2;
    in.pb = b->pb;
    out.coeff_out = b->coeff_out + (size_t)a * 96;
    out.obj = b->obj + a;
    out.status = b->status + a;
    out.iters = b->iters + 2 * a;
    out.lines = b->lines ? b->lines + (size_t)a * 8 * LS * 3 : 0;
    out.line_ok = b->line_ok ? b->line_ok + (size_t)a * 8 * LS : 0;
    int rc = orc_replan(par, &in, &out);
    if (rc)
    {
#ifdef _OPENMP
#pragma omp critical
#endif
      rc_all = rc;
    }
  }
  return rc_all;
}

/*
 * One full replan cycle for one agent, as the reference's planner core runs it around the back end
 * (neptune.cpp:1430-1448 hulls/samples + PredictAlphasBetas, :1512-1529 back end, :719-765
 * post-check against every known trajectory, with the entangle re-check).  Inputs are the
 * committed-trajectory records of all N agents (layout of include/neptune_b200.h) -- this is the
 * CPU baseline that bench.py times next to the device-resident cycle.
 */
#define ORC_REC_TP 16
#define ORC_REC_PWP (1 + (ORC_REC_TP + 1) + 3 * ORC_REC_TP * 4) /* trajectory part of a record */
#define ORC_REC 256 /* record stride: trajectory + DynTraj header (include/neptune_b200.h) */

int orc_cycle_batch(const orc_params* par, int B, const int* agent_id, const int* n_int, const double* coeff_init,
                    const double* t_start, const double* recs, const unsigned char* known, const double* pb,
                    const long long* st_ptr, const double* st_xy, const double* strep, const int* bp_cnt,
                    const double* bp_xy, const int* esv_cnt, const int* esv_alpha, const int* esv_active,
                    const int* es_cnt, const int* es_alpha, const double* es_beta, const int* es_bend,
                    const int* es_active, const double* prev_pos, const double* prev_pos_agent, const double* cur,
                    double delta, int do_entangle, double* coeff_out, double* obj, int* status, int* iters,
                    int* entangled, int* collide, int nthreads, const unsigned char* late, const double* late_recs,
                    const int* bp_cnt_late, const double* bp_xy_late)
{
  const int N = par->num_agents, M = par->num_static, NA = N + M, cap = par->ent_cap, S = par->samples, P = par->num_pol;
  int rc_all = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int b = 0; b < B; b++)
  {
    const unsigned char* kn = known + (size_t)b * N;
    /* hulls and samples of every other agent (neptune.cpp:1433-1434) */
    long long* hptr = (long long*)malloc(sizeof(long long) * ((size_t)N * 8 + 1));
    double* hxy = (double*)malloc(sizeof(double) * 2 * ORC_HMAX * (size_t)N * 8);
    double* nih0 = (double*)malloc(sizeof(double) * (size_t)N * 16);
    double* samp = (double*)calloc((size_t)N * P * (S + 1) * 2, sizeof(double));
    double* samp0 = (double*)calloc((size_t)N * 2, sizeof(double));
    const double d3[3] = { delta, delta, delta };
    long long nv = 0;
    for (int j = 0; j < N; j++)
    {
      const double* rec = recs + (size_t)j * ORC_REC;
      const int np = (int)rec[0];
      const double* times = rec + 1;
      const double* cx = rec + 1 + (ORC_REC_TP + 1);
      const double* cy = cx + ORC_REC_TP * 4;
      for (int i = 0; i < 8; i++)
      {
        hptr[j * 8 + i] = nv;
        nih0[(j * 8 + i) * 2] = nih0[(j * 8 + i) * 2 + 1] = NAN;
        if (!kn[j] || i >= P) continue;
        double h2[2 * ORC_HMAX];
        int hn = 0, h2n = 0, idx[2];
        orc_hull_of_interval(times, np + 1, cx, cy, t_start[b] + i * par->T_span, t_start[b] + (i + 1) * par->T_span,
                             par->T_span, d3, hxy + 2 * nv, &hn, h2, &h2n, idx);
        nv += hn;
        if (h2n > 0)
        {
          nih0[(j * 8 + i) * 2] = h2[0];
          nih0[(j * 8 + i) * 2 + 1] = h2[1];
        }
      }
      if (kn[j])
      {
        orc_sample_interval_points(times, np + 1, cx, cy, t_start[b], t_start[b] + par->T_span * P, P, S,
                                   samp + (size_t)j * P * (S + 1) * 2, 0);
        samp0[2 * j] = samp[(size_t)j * P * (S + 1) * 2];
        samp0[2 * j + 1] = samp[(size_t)j * P * (S + 1) * 2 + 1];
      }
    }
    hptr[N * 8] = nv;
    /* PredictAlphasBetas */
    orc_ent es;
    es.alpha = (int*)malloc(sizeof(int) * 2 * cap);
    es.beta = (double*)malloc(sizeof(double) * cap);
    es.bend = (int*)malloc(sizeof(int) * cap);
    es.active = (int*)malloc(sizeof(int) * NA);
    es.n_alpha = es_cnt[2 * b];
    es.n_bend = es_cnt[2 * b + 1];
    memcpy(es.alpha, es_alpha + (size_t)b * cap * 2, sizeof(int) * 2 * cap);
    memcpy(es.beta, es_beta + (size_t)b * cap, sizeof(double) * cap);
    memcpy(es.bend, es_bend + (size_t)b * cap, sizeof(int) * cap);
    memcpy(es.active, es_active + (size_t)b * NA, sizeof(int) * NA);
    orc_ectx cx;
    cx.N = N, cx.M = M, cx.self = agent_id[b] - 1, cx.cap = cap, cx.pb = pb, cx.strep = strep, cx.bp_cnt = bp_cnt;
    cx.bp_xy = bp_xy, cx.bp_max = par->bp_max;
    int rc = 0;
    if (do_entangle)
      rc = orc_predict(&es, &cx, prev_pos + (size_t)b * (N + 1) * 2, prev_pos_agent + (size_t)b * N * 2, cur + 2 * b,
                       samp0, kn);
    /* back end */
    orc_replan_in in;
    orc_replan_out out;
    in.agent_id = agent_id[b], in.n = n_int[b], in.coeff_init = coeff_init + (size_t)b * 96, in.n_hull_slots = N;
    in.hull_ptr = hptr, in.hull_xy = hxy, in.nih0 = nih0, in.st_ptr = st_ptr, in.st_xy = st_xy;
    in.esv_cnt = esv_cnt + (size_t)b * 18, in.esv_alpha = esv_alpha + (size_t)b * 9 * cap * 2;
    in.esv_active = esv_active + (size_t)b * 9 * NA, in.bp_cnt = bp_cnt, in.bp_xy = bp_xy, in.pb = pb;
    out.coeff_out = coeff_out + (size_t)b * 96, out.obj = obj + b, out.status = status + b, out.iters = iters + 2 * b;
    out.lines = 0, out.line_ok = 0;
    if (!rc) rc = orc_replan(par, &in, &out);
    /* safetyCheckAfterReplan (neptune.cpp:719-752) against the trajectories that arrived during the optimisation
     * (late == NULL: every known one, with the record it was planned against) */
    const unsigned char* lt = late ? late + (size_t)b * N : kn;
    const double* lrecs = late_recs ? late_recs : recs;
    int col = 0, any_late = 0;
    for (int j = 0; j < N; j++)
      if (lt[j] && j != agent_id[b] - 1)
      {
        any_late = 1;
        if (col) continue;
        const double* rec = lrecs + (size_t)j * ORC_REC;
        const double* cxj = rec + 1 + (ORC_REC_TP + 1);
        if (orc_pwp_collides(out.coeff_out, n_int[b], t_start[b], par->T_span, rec + 1, (int)rec[0] + 1, cxj,
                             cxj + ORC_REC_TP * 4, d3))
          col = 1;
      }
    collide[b] = col;
    entangled[b] = 0;
    if (do_entangle && !rc && any_late)
    { /* :735-752: samples of the late agents re-drawn over the optimised trajectory's span, their bend points from the
         late message, PredictAlphasBetas afresh from entangle_state_, then entangleCheckGivenPwp */
      unsigned char* kn2 = (unsigned char*)malloc((size_t)N);
      int* bc2 = (int*)malloc(sizeof(int) * (size_t)N);
      double* bx2 = (double*)malloc(sizeof(double) * (size_t)N * par->bp_max * 2);
      memcpy(bc2, bp_cnt, sizeof(int) * (size_t)N);
      memcpy(bx2, bp_xy, sizeof(double) * (size_t)N * par->bp_max * 2);
      for (int j = 0; j < N; j++)
      {
        kn2[j] = kn[j];
        if (!lt[j] || j == agent_id[b] - 1) continue;
        kn2[j] = 1;
        const double* rec = lrecs + (size_t)j * ORC_REC;
        const double* cxj = rec + 1 + (ORC_REC_TP + 1);
        orc_sample_interval_points(rec + 1, (int)rec[0] + 1, cxj, cxj + ORC_REC_TP * 4, t_start[b],
                                   t_start[b] + par->T_span * n_int[b], P, S, samp + (size_t)j * P * (S + 1) * 2, 0);
        samp0[2 * j] = samp[(size_t)j * P * (S + 1) * 2];
        samp0[2 * j + 1] = samp[(size_t)j * P * (S + 1) * 2 + 1];
        if (bp_cnt_late)
        {
          bc2[j] = bp_cnt_late[j];
          memcpy(bx2 + (size_t)j * par->bp_max * 2, bp_xy_late + (size_t)j * par->bp_max * 2, sizeof(double) * par->bp_max * 2);
        }
      }
      es.n_alpha = es_cnt[2 * b];
      es.n_bend = es_cnt[2 * b + 1];
      memcpy(es.alpha, es_alpha + (size_t)b * cap * 2, sizeof(int) * 2 * cap);
      memcpy(es.beta, es_beta + (size_t)b * cap, sizeof(double) * cap);
      memcpy(es.bend, es_bend + (size_t)b * cap, sizeof(int) * cap);
      memcpy(es.active, es_active + (size_t)b * NA, sizeof(int) * NA);
      orc_ectx cx2 = cx;
      cx2.bp_cnt = bc2, cx2.bp_xy = bx2;
      int e = orc_predict(&es, &cx2, prev_pos + (size_t)b * (N + 1) * 2, prev_pos_agent + (size_t)b * N * 2, cur + 2 * b, samp0, kn2);
      if (!e)
      {
        double cxy[2 * 32];
        for (int i = 0; i < n_int[b]; i++)
          for (int r = 0; r < 4; r++)
          {
            cxy[4 * i + r] = out.coeff_out[4 * i + r];
            cxy[4 * n_int[b] + 4 * i + r] = out.coeff_out[32 + 4 * i + r];
          }
        e = orc_entangle_check_pwp(&es, &cx2, n_int[b], cxy, samp, kn2, P, S, par->T_span);
      }
      entangled[b] = e > 0;
      if (e < 0) rc = e;
      free(kn2), free(bc2), free(bx2);
    }
    free(hptr), free(hxy), free(nih0), free(samp), free(samp0), free(es.alpha), free(es.beta), free(es.bend), free(es.active);
    if (rc)
    {
#ifdef _OPENMP
#pragma omp critical
#endif
      rc_all = rc;
    }
  }
  return rc_all;
}

/*
 * mu::composePieceWisePol (utils.cpp:318-402) on committed-trajectory records (layout of
 * include/neptune_b200.h): out = the pieces of p1 that start after t and before p2 begins, then p2
 * (caller: Neptune::replanFull neptune.cpp:1689-1699).  p1 and p2 are modified in place exactly as the
 * reference modifies its arguments (times.front() adjustments :320-336).  Returns the number of pieces
 * of out, 0 for the "dummy" empty result (:342-354), -1 if out would exceed 16 pieces.
 */
int orc_compose_records(double t, double dc, double* p1, double* p2, double* out)
{
  (void)dc;
  int n1 = (int)p1[0], n2 = (int)p2[0];
  double* t1 = p1 + 1;
  double* t2 = p2 + 1;
  if (n1 < 1 || n2 < 1 || n1 > ORC_REC_TP || n2 > ORC_REC_TP) /* empty pwp: the reference reads .back() of an empty vector */
  {
    memset(out, 0, sizeof(double) * ORC_REC);
    return 0;
  }
  if (t > t1[n1] && t < t2[0]) t2[0] = t;
  if (t1[n1] < t2[0]) t2[0] = t1[n1];
  if (t < t1[0]) t1[0] = t;
  if (fabs(t - t2[0]) < 1e-5)
  {
    memcpy(out, p2, sizeof(double) * ORC_REC);
    return n2;
  }
  memset(out, 0, sizeof(double) * ORC_REC);
  if (t1[n1] < t2[0] || t > t2[n2] || t < t1[0]) return 0;
  double* to = out + 1;
  int np = 0;
  to[0] = t;
#define COPY_PIECE(src, k)                                                                   \
  do                                                                                         \
  {                                                                                          \
    if (np >= ORC_REC_TP) return -1;                                                         \
    for (int ax = 0; ax < 3; ax++)                                                           \
      memcpy(out + 1 + (ORC_REC_TP + 1) + ax * ORC_REC_TP * 4 + 4 * np,                      \
             (src) + 1 + (ORC_REC_TP + 1) + ax * ORC_REC_TP * 4 + 4 * (k), sizeof(double) * 4); \
  } while (0)
  for (int i = 1; i <= n1; i++) /* i = 0 never qualifies: t1[0] <= t after :332-336 */
    if (t1[i] > t && t1[i] < t2[0])
    {
      COPY_PIECE(p1, i - 1);
      np++;
      to[np] = t1[i];
    }
  for (int i = 0; i <= n2; i++)
    if (t2[i] > t)
    {
      if (i == 0)
        COPY_PIECE(p1, n1 - 1);
      else
        COPY_PIECE(p2, i - 1);
      np++;
      to[np] = t2[i];
    }
#undef COPY_PIECE
  out[0] = (double)np;
  memcpy(out + ORC_REC_PWP, p2 + ORC_REC_PWP, sizeof(double) * (ORC_REC - ORC_REC_PWP)); /* header of the new publication */
  return np;
}

/*
 * Tail of Neptune::replanFull (neptune.cpp:1685-1699) for a batch: pwp_now = coefficients with times shifted
 * by t_start (generatePwpOut, solver_gurobi_poly.cpp:892-907), pwp_out = composePieceWisePol(time_now, dc,
 * pwp_prev, pwp_now).  A rejected replan (status >= 2, entangled, collide) keeps the previous record.
 * recs [N][210] (prev of agent b = recs[agent_id[b] - 1]); new_recs [B][210]; n_pieces [B].
 */
int orc_commit_compose_batch(const orc_params* par, int B, const int* agent_id, const int* n_int, const double* coeff_out,
                             const double* t_start, const double* t_now, const double* recs, const int* status,
                             const int* entangled, const int* collide, double* new_recs, int* n_pieces)
{
  int rc = 0;
  for (int b = 0; b < B; b++)
  {
    double now[ORC_REC], prev[ORC_REC];
    double* out = new_recs + (size_t)b * ORC_REC;
    memset(now, 0, sizeof(now));
    memcpy(prev, recs + (size_t)(agent_id[b] - 1) * ORC_REC, sizeof(prev));
    const int n = n_int[b];
    now[0] = (double)n;
    for (int k = 0; k <= n; k++) now[1 + k] = t_start[b] + (double)k * par->T_span;
    for (int ax = 0; ax < 3; ax++)
      for (int i = 0; i < n; i++)
        for (int c = 0; c < 4; c++)
          now[1 + (ORC_REC_TP + 1) + ax * ORC_REC_TP * 4 + 4 * i + c] = coeff_out[(size_t)b * 96 + ax * 32 + 4 * i + c];
    if (status[b] >= 2 || entangled[b] || collide[b])
    {
      memcpy(out, prev, sizeof(prev));
      n_pieces[b] = (int)prev[0];
      continue;
    }
    int np = orc_compose_records(t_now[b], 0.0 /* dc: unused by the reference body */, prev, now, out);
    if (np < 0) rc = -1, np = 0;
    n_pieces[b] = np;
  }
  return rc;
}
