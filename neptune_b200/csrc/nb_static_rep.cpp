// nb_static_rep.cpp -- host-side, one-time set-up of the static-obstacle representation the entanglement chain uses.
//
// Replaces NeptuneRos::setUpCheckingPosAndStaticObs (reference neptune/src/neptune_ros.cpp:852-1019): every static
// obstacle is reduced to the two points where a line through its centre leaves it; the common slope of those lines is
// found by sweeping one degree at a time from the base -> position direction until (a) the parallel lines through
// any two obstacle centres, and through the agent's base, are at least 1.6 voxels apart, (b) no line cuts the
// straight tether base -> position, (c) each line leaves its polygon on both sides of the centre.  Also returns
// staticObsLongestDist (:986-1002).  Pure host code like the reference's (it runs once per agent at start-up); the
// results are uploaded with nb_set_static / nb_set_static_longest.
#include <cmath>
#include <vector>

#include "../../include/neptune_b200.h"

namespace
{
struct Pt
{
  double x, y;
};

// parameter along c -> d where the line a -> b meets it, and the determinant (:913-921, :946-953)
inline double meet(const Pt& a, const Pt& b, const Pt& c, const Pt& d, double* det)
{
  *det = (b.y - a.y) * (d.x - c.x) - (b.x - a.x) * (d.y - c.y);
  return ((c.y - a.y) * (b.x - a.x) - (b.y - a.y) * (c.x - a.x)) / *det;
}
}  // namespace

extern "C" int nb_static_obst_rep(int32_t M, const int64_t* poly_ptr, const double* poly_xy, const double* base,
                                  const double* pos, double voxel_size, double* strep, double* longest)
{
  if (M < 0 || (M > 0 && (!poly_ptr || !poly_xy || !strep || !longest)) || !base || !pos) return NB_ERR_ARG;
  std::vector<Pt> centre(M);
  for (int i = 0; i < M; i++)
  {
    const int nv = (int)(poly_ptr[i + 1] - poly_ptr[i]);
    if (nv < 1) return NB_ERR_ARG;
    Pt s{ 0.0, 0.0 };
    for (int j = 0; j < nv; j++) s.x += poly_xy[2 * (poly_ptr[i] + j)], s.y += poly_xy[2 * (poly_ptr[i] + j) + 1];
    centre[i] = Pt{ s.x / nv, s.y / nv };
  }
  const Pt pb{ base[0], base[1] }, p0{ pos[0], pos[1] };
  const double theta0 = atan2(p0.y - pb.y, p0.x - pb.x);
  const double min_gap = voxel_size * 1.6;
  for (double theta = theta0;; theta = theta + 1.0 / 180.0 * 3.1415927)
  {
    if (theta > theta0 + 3.14) return NB_ERR_ARG;  // "cannot find a feasible vertex representation" (:882-887)
    const double m = tan(theta);
    const double norm = sqrt(1 + m * m);
    bool good = true;
    for (int i = 0; i < M && good; i++)
    {  // (a) spacing of the parallel lines, (b) no cut through the straight tether
      const double e1 = centre[i].y - centre[i].x * m;
      for (int j = i + 1; j < M && good; j++)
        if (fabs(e1 - (centre[j].y - centre[j].x * m)) / norm < min_gap) good = false;
      if (good && fabs(e1 - (pb.y - pb.x * m)) / norm < min_gap) good = false;
      if (!good) break;
      double det;
      const double t = meet(centre[i], Pt{ centre[i].x + 10, centre[i].y + 10 * m }, pb, p0, &det);
      if (!(fabs(det) < 0.01 || t < 0 || t > 1.0)) good = false;
    }
    for (int i = 0; i < M && good; i++)
    {  // (c) exit points on both sides of the centre
      const int nv = (int)(poly_ptr[i + 1] - poly_ptr[i]);
      const double* v = poly_xy + 2 * poly_ptr[i];
      const Pt a = centre[i], b{ centre[i].x + 10, centre[i].y + 10 * m };
      double hi = -1e-5, lo = 1e-5;
      Pt vhi{ 0, 0 }, vlo{ 0, 0 };
      for (int j = 0; j < nv; j++)
      {
        const Pt c{ v[2 * j], v[2 * j + 1] }, d{ v[2 * ((j + 1) % nv)], v[2 * ((j + 1) % nv) + 1] };
        double det;
        const double t = meet(a, b, c, d, &det);
        if (fabs(det) < 0.01 || t < 0 || t > 1.0) continue;
        const double along = fabs(b.x - a.x) < 0.01 ? (c.y - a.y) / (b.y - a.y) + t * (d.y - c.y) / (b.y - a.y)
                                                    : (c.x - a.x) / (b.x - a.x) + t * (d.x - c.x) / (b.x - a.x);
        if (along > hi)
          hi = along, vhi = Pt{ c.x + t * (d.x - c.x), c.y + t * (d.y - c.y) };
        else if (along < lo)
          lo = along, vlo = Pt{ c.x + t * (d.x - c.x), c.y + t * (d.y - c.y) };
      }
      if (hi < 0 || lo > 0)
      {
        good = false;
        break;
      }
      strep[4 * i] = vlo.x, strep[4 * i + 1] = vlo.y, strep[4 * i + 2] = vhi.x, strep[4 * i + 3] = vhi.y;
      double far_lo = 0, far_hi = 0;
      for (int j = 0; j < nv; j++)
      {
        const double dl = sqrt((v[2 * j] - vlo.x) * (v[2 * j] - vlo.x) + (v[2 * j + 1] - vlo.y) * (v[2 * j + 1] - vlo.y));
        const double dh = sqrt((v[2 * j] - vhi.x) * (v[2 * j] - vhi.x) + (v[2 * j + 1] - vhi.y) * (v[2 * j + 1] - vhi.y));
        if (dl > far_lo) far_lo = dl;
        if (dh > far_hi) far_hi = dh;
      }
      longest[2 * i] = far_lo, longest[2 * i + 1] = far_hi;
    }
    if (good) return NB_OK;
  }
}
