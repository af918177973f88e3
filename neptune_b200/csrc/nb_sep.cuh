// nb_sep.cuh -- K2: separating-line LP as a closest-pair reduction.
//
// Replaces separator::Separator::solveModel, 2-D overloads (reference
// submodules/separator/src/separator_glpk.cpp:248-373, :375-498).  The LP there is
//     find (n0,n1,d):  n.a + d >= 1  for a in A (u A+),   n.b + d <= -1  for b in B,  objective 0,
// and GLPK returns an implementation-defined feasible vertex.  This kernel returns the CANONICAL
// feasible point: the minimum-norm (maximum-margin) solution.  With pa in conv(A), pb in conv(B) the
// closest pair and delta = |pa - pb| > 0, it is n = 2 (pa - pb) / delta^2, d = -n.(pa + pb)/2
// (value +1 at pa, -1 at pb).  The LP is feasible iff that line separates, which is verified row by
// row with the slack NB_SEP_EPS the oracle uses too.  The slack is 1e-6 of the (unit) margin, not 1e-9: for sets that
// nearly touch (delta ~ 1e-3) along nearly parallel edges, n = 2 (pa - pb) / delta^2 is a small difference of large
// coordinates scaled up by 1 / delta^2, and the rounding of its DIRECTION alone moves the row value of a vertex a few
// metres along the edge by ~1e-8: with 1e-9 the solved flag of such a pair flipped with a 1e-15 perturbation of the
// control points (the row would then be dropped, where GLPK -- feasibility tolerance 1e-7 -- reports the LP feasible),
// and the oracle's enumeration fell back to a candidate of smaller margin.  Intersecting sets fail the check by O(1).
#pragma once
#include "nb_common.cuh"

// squared distance from p to segment (u,v); c = closest point on the segment
NB_HD double nb_pt_seg(double px, double py, double ux, double uy, double vx, double vy, double& cx, double& cy)
{
  const double ex = vx - ux, ey = vy - uy;
  const double e2 = ex * ex + ey * ey;
  double t = 0.0;
  if (e2 > 0.0)
  {
    t = ((px - ux) * ex + (py - uy) * ey) / e2;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  }
  cx = ux + t * ex;
  cy = uy + t * ey;
  const double dx = px - cx, dy = py - cy;
  return dx * dx + dy * dy;
}

NB_HD double nb_sep_rcp(double x)
{
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// squared distance from p to the segment u + t e, t in [0,1], with inv = 1/|e|^2 (0 for a degenerate segment)
NB_HD double nb_pt_seg_inv(double px, double py, double ux, double uy, double ex, double ey, double inv, double& cx,
                           double& cy)
{
  double t = ((px - ux) * ex + (py - uy) * ey) * inv;
  t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
  cx = ux + t * ex;
  cy = uy + t * ey;
  const double dx = px - cx, dy = py - cy;
  return dx * dx + dy * dy;
}

// A: nA points (x,y interleaved); a_polygon: vertices are in cyclic order of a convex polygon.
// B: nB points (the agent's MINVO control points; any order).  Returns true iff separable.
// Reciprocals of the squared segment lengths are hoisted: 6 for the segments of B (nB = 4) and one per
// edge of A, instead of one division per point-segment pair.
NB_HD bool nb_separate(const double* A, int nA, bool a_polygon, const double* B, int nB, double out[3])
{
  double best = 1e300, pax = 0, pay = 0, pbx = 0, pby = 0, cx, cy;
  out[0] = out[1] = out[2] = 0.0;
  if (nA <= 0 || nB <= 0) return false;
  // segments of B (all pairs; a single point is its own degenerate segment)
  double sux[6], suy[6], sex[6], sey[6], sinv[6];
  int ns = 0;
  if (nB <= 4)
  {
    for (int j = 0; j < nB; j++)
      for (int k = (nB == 1 ? j : j + 1); k < nB; k++)
      {
        sux[ns] = B[2 * j], suy[ns] = B[2 * j + 1];
        sex[ns] = B[2 * k] - B[2 * j], sey[ns] = B[2 * k + 1] - B[2 * j + 1];
        const double e2 = sex[ns] * sex[ns] + sey[ns] * sey[ns];
        sinv[ns] = e2 > 0.0 ? nb_sep_rcp(e2) : 0.0;
        ns++;
      }
    for (int i = 0; i < nA; i++)
    {
      const double ax = A[2 * i], ay = A[2 * i + 1];
      for (int q = 0; q < ns; q++)
      {
        const double d2 = nb_pt_seg_inv(ax, ay, sux[q], suy[q], sex[q], sey[q], sinv[q], cx, cy);
        if (d2 < best)
        {
          best = d2;
          pax = ax, pay = ay, pbx = cx, pby = cy;
        }
      }
    }
  }
  else
  {
    for (int i = 0; i < nA; i++)
    {
      const double ax = A[2 * i], ay = A[2 * i + 1];
      for (int j = 0; j < nB; j++)
        for (int k = j + 1; k < nB; k++)
        {
          const double d2 = nb_pt_seg(ax, ay, B[2 * j], B[2 * j + 1], B[2 * k], B[2 * k + 1], cx, cy);
          if (d2 < best)
          {
            best = d2;
            pax = ax, pay = ay, pbx = cx, pby = cy;
          }
        }
    }
  }
  // vertices of B against the segments of A: polygon edges when ordered, else all pairs
  if (a_polygon || nA <= 2)
  {
    const int ne = nA <= 2 ? 1 : nA;
    for (int j = 0; j < ne; j++)
    {
      const int k = (j + 1 < nA) ? j + 1 : 0;
      const double ux = A[2 * j], uy = A[2 * j + 1], ex = A[2 * k] - ux, ey = A[2 * k + 1] - uy;
      const double e2 = ex * ex + ey * ey;
      const double inv = e2 > 0.0 ? nb_sep_rcp(e2) : 0.0;
      for (int i = 0; i < nB; i++)
      {
        const double bx = B[2 * i], by = B[2 * i + 1];
        const double d2 = nb_pt_seg_inv(bx, by, ux, uy, ex, ey, inv, cx, cy);
        if (d2 < best)
        {
          best = d2;
          pax = cx, pay = cy, pbx = bx, pby = by;
        }
      }
    }
  }
  else
  {
    for (int j = 0; j < nA; j++)
      for (int k = j + 1; k < nA; k++)
      {
        const double ux = A[2 * j], uy = A[2 * j + 1], ex = A[2 * k] - ux, ey = A[2 * k + 1] - uy;
        const double e2 = ex * ex + ey * ey;
        const double inv = e2 > 0.0 ? nb_sep_rcp(e2) : 0.0;
        for (int i = 0; i < nB; i++)
        {
          const double bx = B[2 * i], by = B[2 * i + 1];
          const double d2 = nb_pt_seg_inv(bx, by, ux, uy, ex, ey, inv, cx, cy);
          if (d2 < best)
          {
            best = d2;
            pax = cx, pay = cy, pbx = bx, pby = by;
          }
        }
      }
  }
  if (!(best > 1e-24)) return false;
  const double ux = pax - pbx, uy = pay - pby;
  const double ib = 2.0 * nb_sep_rcp(best);
  const double n0 = ux * ib, n1 = uy * ib;
  const double d = -(n0 * (pax + pbx) + n1 * (pay + pby)) * 0.5;
  for (int i = 0; i < nA; i++)
    if (!(n0 * A[2 * i] + n1 * A[2 * i + 1] + d >= 1.0 - NB_SEP_EPS)) return false;
  for (int i = 0; i < nB; i++)
    if (!(n0 * B[2 * i] + n1 * B[2 * i + 1] + d <= -1.0 + NB_SEP_EPS)) return false;
  out[0] = n0;
  out[1] = n1;
  out[2] = d;
  return true;
}
