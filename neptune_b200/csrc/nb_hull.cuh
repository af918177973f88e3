// nb_hull.cuh -- K1 (hulls and samples of the other agents' committed trajectories) and K5 (GJK
// post-check).
//
// K1 replaces Neptune::convexHullsOfCurves2d / convexHullsOfCurve2d / convexHullOfInterval2d /
// vertexesOfInterval2d (reference neptune/src/neptune.cpp:224-267, :269-285, :288-325, :349-452) with
// cu::convexHullOfPoints2d (cgal_utils.cpp:157-174), and Neptune::SamplePointsOfCurves /
// SamplePointsOfIntervals (neptune.cpp:463-498, :500-566).  One thread per (planning agent, other
// agent, window).  The hull is Andrew's monotone chain on the same operands and in the same operation
// order as the oracle (no FMA contraction), so vertex sets, vertex order and interval indices are
// bit-exact.  CGAL convention adopted: strict extreme points, counter-clockwise from the
// lexicographically smallest point.
//
// K5 replaces Neptune::trajsAndPwpAreInCollision2d (neptune.cpp:767-806) with gjk::collision
// (gjk.cpp:76-148), the geometric half of Neptune::safetyCheckAfterReplan (neptune.cpp:719-765).
#pragma once
#include "nb_common.cuh"

#define NB_TP 16                           // pieces a committed trajectory can hold
#define NB_REC_PWP (1 + (NB_TP + 1) + 3 * NB_TP * 4)  // doubles of the trajectory part of a record (210)
#define NB_REC 256                         // doubles per committed-trajectory record (2 KB): trajectory + DynTraj header
// DynTraj header of a record (mader_msgs/msg/DynTraj.msg:3-9), after the trajectory part
#define NB_REC_ID 210       // id (1-based)
#define NB_REC_ISAGENT 211  // is_agent
#define NB_REC_BBOX 212     // bbox[3]
#define NB_REC_POS 215      // pos[3]: position of the publisher when it published
#define NB_REC_NBEND 218    // bendpt.size() (tether base included)
#define NB_REC_BEND 219     // bendpt[8][2]
#define NB_REC_BEND_MAX 8
#define NB_REC_SEQ 235      // cycle in which the record was committed (stands in for time_received)
#define NB_HMAX 24                         // vertices of an inflated hull (<= 4 + control points)
#define NB_HPCS 4                          // pieces one window may overlap

// committed-trajectory record (the payload of the per-cycle exchange): [0] n_pieces, [1..17] times, coeff[3][16][4],
// then the DynTraj header above
NB_HD int nb_rec_np(const double* rec) { return (int)rec[0]; }
NB_HD const double* nb_rec_times(const double* rec) { return rec + 1; }
NB_HD const double* nb_rec_coeff(const double* rec, int ax) { return rec + 1 + (NB_TP + 1) + ax * NB_TP * 4; }

NB_HD double nb_cross3(const double* o, const double* a, const double* b)
{
  return NB_SUB(NB_MUL(NB_SUB(a[0], o[0]), NB_SUB(b[1], o[1])), NB_MUL(NB_SUB(a[1], o[1]), NB_SUB(b[0], o[0])));
}

NB_HD bool nb_pt_less(const double* p, const double* q) { return p[0] < q[0] || (p[0] == q[0] && p[1] < q[1]); }

// strict convex hull of n points (destroys pts: sorted in place), CCW from the lexicographic minimum
NB_HD int nb_convex_hull(double* pts, int n, double* out, int out_cap)
{
  for (int i = 1; i < n; i++)
  {  // insertion sort, lexicographic
    const double x = pts[2 * i], y = pts[2 * i + 1];
    int j = i - 1;
    while (j >= 0 && (x < pts[2 * j] || (x == pts[2 * j] && y < pts[2 * j + 1])))
    {
      pts[2 * j + 2] = pts[2 * j];
      pts[2 * j + 3] = pts[2 * j + 1];
      j--;
    }
    pts[2 * j + 2] = x;
    pts[2 * j + 3] = y;
  }
  int m = 0;
  for (int i = 0; i < n; i++)
    if (m == 0 || pts[2 * i] != pts[2 * (m - 1)] || pts[2 * i + 1] != pts[2 * (m - 1) + 1])
    {
      pts[2 * m] = pts[2 * i];
      pts[2 * m + 1] = pts[2 * i + 1];
      m++;
    }
  if (m <= 2)
  {
    for (int i = 0; i < m && i < out_cap; i++) out[2 * i] = pts[2 * i], out[2 * i + 1] = pts[2 * i + 1];
    return m;
  }
  // the chain is built in a scratch that can hold every point: reuse a local buffer
  double h[2 * (4 * 4 * NB_HPCS + 2)];
  int k = 0;
  for (int i = 0; i < m; i++)
  {
    while (k >= 2 && nb_cross3(h + 2 * (k - 2), h + 2 * (k - 1), pts + 2 * i) <= 0) k--;
    h[2 * k] = pts[2 * i], h[2 * k + 1] = pts[2 * i + 1];
    k++;
  }
  for (int i = m - 2, t = k + 1; i >= 0; i--)
  {
    while (k >= t && nb_cross3(h + 2 * (k - 2), h + 2 * (k - 1), pts + 2 * i) <= 0) k--;
    h[2 * k] = pts[2 * i], h[2 * k + 1] = pts[2 * i + 1];
    k++;
  }
  k--;
  const int kk = k < out_cap ? k : out_cap;
  for (int i = 0; i < kk; i++) out[2 * i] = h[2 * i], out[2 * i + 1] = h[2 * i + 1];
  return k;
}

NB_HD int nb_lower_bound(const double* a, int n, double v)
{
  int i = 0;
  while (i < n && a[i] < v) i++;
  return i;
}
NB_HD int nb_upper_bound(const double* a, int n, double v)
{
  int i = 0;
  while (i < n && !(v < a[i])) i++;
  return i;
}
NB_HD int nb_sat(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// MINVO control points of every piece overlapping [t0,t1] (neptune.cpp:379-448).  cps: [<=16][2].
// idx = {index_first_interval, index_last_interval}.  Returns the number of control points, or -1
// when the window overlaps more than NB_HPCS pieces.
NB_HD int nb_window_ctrl_pts(const NbConsts& cs, const double* rec, double t0, double t1, double* cps, int* idx)
{
  const int np = nb_rec_np(rec), nt = np + 1;
  const double* times = nb_rec_times(rec);
  const double* cx = nb_rec_coeff(rec, 0);
  const double* cy = nb_rec_coeff(rec, 1);
  int first = nb_sat(nb_lower_bound(times, nt, t0) - 1, 0, np - 1);
  int last = nb_sat(nb_upper_bound(times, nt, t1) - 1, 0, np - 1);
  idx[0] = first;
  idx[1] = last;
  if (last - first + 1 > NB_HPCS) return -1;
  int n = 0;
  for (int i = first; i <= last; i++)
  {
    double t;
    if (i != last)
      t = NB_SUB(times[i + 1], times[i]);
    else if (t1 > times[i + 1])
      t = NB_SUB(times[i + 1], times[i]);
    else
      t = NB_SUB(t1, times[i]);
    if (t > cs.T)
      t = cs.T;
    else if (t < 0)
      t = 0;
    const double sc[4] = { NB_MUL(NB_MUL(t, t), t), NB_MUL(t, t), t, 1.0 };
    for (int k = 0; k < 4; k++)
    {
      double x = 0, y = 0;
      for (int r = 0; r < 4; r++)
      {
        x = NB_ADD(x, NB_MUL(NB_MUL(cx[4 * i + r], sc[r]), cs.Ainv01[r * 4 + k]));
        y = NB_ADD(y, NB_MUL(NB_MUL(cy[4 * i + r], sc[r]), cs.Ainv01[r * 4 + k]));
      }
      cps[2 * n] = x, cps[2 * n + 1] = y;
      n++;
    }
  }
  return n;
}

// monotone chain over points that are ALREADY sorted lexicographically (duplicates allowed)
NB_HD int nb_chain_sorted(double* pts, int n, double* out, int out_cap)
{
  int m = 0;
  for (int i = 0; i < n; i++)
    if (m == 0 || pts[2 * i] != pts[2 * (m - 1)] || pts[2 * i + 1] != pts[2 * (m - 1) + 1])
    {
      pts[2 * m] = pts[2 * i];
      pts[2 * m + 1] = pts[2 * i + 1];
      m++;
    }
  if (m <= 2)
  {
    for (int i = 0; i < m && i < out_cap; i++) out[2 * i] = pts[2 * i], out[2 * i + 1] = pts[2 * i + 1];
    return m;
  }
  double h[2 * (4 * 4 * NB_HPCS + 2)];
  int k = 0;
  for (int i = 0; i < m; i++)
  {
    while (k >= 2 && nb_cross3(h + 2 * (k - 2), h + 2 * (k - 1), pts + 2 * i) <= 0) k--;
    h[2 * k] = pts[2 * i], h[2 * k + 1] = pts[2 * i + 1];
    k++;
  }
  for (int i = m - 2, t = k + 1; i >= 0; i--)
  {
    while (k >= t && nb_cross3(h + 2 * (k - 2), h + 2 * (k - 1), pts + 2 * i) <= 0) k--;
    h[2 * k] = pts[2 * i], h[2 * k + 1] = pts[2 * i + 1];
    k++;
  }
  k--;
  const int kk = k < out_cap ? k : out_cap;
  for (int i = 0; i < kk; i++) out[2 * i] = h[2 * i], out[2 * i + 1] = h[2 * i + 1];
  return k;
}

// Neptune::convexHullOfInterval2d (neptune.cpp:288-309): inflated hull + col(0) of the un-inflated one.
// The 4 x nc inflated corners are sorted by sorting the nc control points once and merging the four
// translated copies (a translation keeps the lexicographic order): the same sorted sequence the plain
// sort gives, hence the same hull, at a quarter of the work.
NB_HD int nb_hull_of_window(const NbConsts& cs, const double* rec, double t0, double t1, double delta, double* hull,
                            double* nih0, int* idx)
{
  double cps[2 * 4 * NB_HPCS], pts[2 * 16 * NB_HPCS];
  const int nc = nb_window_ctrl_pts(cs, rec, t0, t1, cps, idx);
  if (nc < 0) return -1;
  for (int i = 1; i < nc; i++)
  {  // insertion sort of the control points, lexicographic
    const double x = cps[2 * i], y = cps[2 * i + 1];
    int j = i - 1;
    while (j >= 0 && (x < cps[2 * j] || (x == cps[2 * j] && y < cps[2 * j + 1])))
    {
      cps[2 * j + 2] = cps[2 * j];
      cps[2 * j + 3] = cps[2 * j + 1];
      j--;
    }
    cps[2 * j + 2] = x;
    cps[2 * j + 3] = y;
  }
  nih0[0] = cps[0];  // first vertex of a CCW-from-lexicographic-minimum hull
  nih0[1] = cps[1];
  // groups in lexicographic order of their offsets: (-,-) (-,+) (+,-) (+,+)
  const double ox[4] = { -1, -1, 1, 1 }, oy[4] = { -1, 1, -1, 1 };
  int head[4] = { 0, 0, 0, 0 };
  int np = 0;
  for (int q = 0; q < 4 * nc; q++)
  {
    int bg = -1;
    double bx = 0, by = 0;
    for (int gI = 0; gI < 4; gI++)
    {
      if (head[gI] >= nc) continue;
      const double x = NB_ADD(cps[2 * head[gI]], NB_MUL(ox[gI], delta)), y = NB_ADD(cps[2 * head[gI] + 1], NB_MUL(oy[gI], delta));
      if (bg < 0 || x < bx || (x == bx && y < by))
      {
        bg = gI;
        bx = x;
        by = y;
      }
    }
    head[bg]++;
    pts[2 * np] = bx, pts[2 * np + 1] = by;
    np++;
  }
  return nb_chain_sorted(pts, np, hull, NB_HMAX);
}

#if defined(__CUDACC__)
// ---- nb_hull_of_window by ONE WARP (k_hulls): the same points, the same sorted sequence and the same chain as the
// scalar version above, hence the same hull bit for bit, but nothing sequential except the chain itself:
//   * piece range by ballot over the knot vector, control points one lane each;
//   * the 4 nc inflated corners one or two per lane, sorted by RANK: a corner's place is the number of corners that
//     precede it lexicographically (ties by index; equal corners are dropped afterwards, so their order is immaterial);
//   * lower and upper chain at the same time on lanes 0 and 1 (Andrew's two chains do not depend on each other; the
//     scalar code's `t = k + 1` floor for the second loop is exactly "the upper chain starts at the last point").
// scratch: NB_HULL_WARP_SCRATCH doubles of shared memory per warp.
#define NB_HULL_WARP_SCRATCH (2 * 64 + 2 * 64 + 2 * 2 * 66)
__device__ __forceinline__ int nb_hull_of_window_warp(const NbConsts& cs, const double* rec, double t0, double t1, double delta,
                                                      double* hull /*global [NB_HMAX][2]*/, double* nih0, int* idx, double* scr,
                                                      int lane)
{
  const unsigned FULL = 0xffffffffu;
  double* pts = scr;             // [64][2] inflated corners, unsorted
  double* srt = scr + 128;       // [64][2] sorted, then de-duplicated in place
  double* chn = scr + 256;       // [2][66][2] lower / upper chain
  const int np = nb_rec_np(rec), nt = np + 1;
  const double* times = nb_rec_times(rec);
  const double* cx = nb_rec_coeff(rec, 0);
  const double* cy = nb_rec_coeff(rec, 1);
  const double tl = lane < nt ? times[lane] : 0.0;
  // lower_bound(t0): first i with !(times[i] < t0); upper_bound(t1): first i with t1 < times[i]
  const unsigned in = nt >= 32 ? FULL : ((1u << nt) - 1u);
  const unsigned nlt = ~__ballot_sync(FULL, lane < nt && tl < t0) & in, gt = __ballot_sync(FULL, lane < nt && t1 < tl);
  const int lb = nlt ? __ffs(nlt) - 1 : nt, ub = gt ? __ffs(gt) - 1 : nt;
  const int first = nb_sat(lb - 1, 0, np - 1), last = nb_sat(ub - 1, 0, np - 1);
  idx[0] = first, idx[1] = last;
  if (last - first + 1 > NB_HPCS) return -1;
  const int nc = 4 * (last - first + 1);
  // control point `lane` (piece first + lane / 4, MINVO column lane % 4): neptune.cpp:379-448, same operation order
  double px = 0, py = 0;
  if (lane < nc)
  {
    const int i = first + (lane >> 2), k = lane & 3;
    double t;
    if (i != last)
      t = NB_SUB(times[i + 1], times[i]);
    else if (t1 > times[i + 1])
      t = NB_SUB(times[i + 1], times[i]);
    else
      t = NB_SUB(t1, times[i]);
    if (t > cs.T)
      t = cs.T;
    else if (t < 0)
      t = 0;
    const double sc[4] = { NB_MUL(NB_MUL(t, t), t), NB_MUL(t, t), t, 1.0 };
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
      px = NB_ADD(px, NB_MUL(NB_MUL(cx[4 * i + r], sc[r]), cs.Ainv01[r * 4 + k]));
      py = NB_ADD(py, NB_MUL(NB_MUL(cy[4 * i + r], sc[r]), cs.Ainv01[r * 4 + k]));
    }
  }
  {  // col(0) of the un-inflated hull: the lexicographic minimum of the control points
    double mx = lane < nc ? px : __longlong_as_double(0x7ff0000000000000LL), my = lane < nc ? py : mx;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
    {
      const double ox = __shfl_xor_sync(FULL, mx, o), oy = __shfl_xor_sync(FULL, my, o);
      if (ox < mx || (ox == mx && oy < my)) mx = ox, my = oy;
    }
    nih0[0] = __shfl_sync(FULL, mx, 0), nih0[1] = __shfl_sync(FULL, my, 0);
  }
  // inflated corners: corner p = 4 c + g is control point c moved by the g-th of (-,-) (-,+) (+,-) (+,+)
  const int n4 = 4 * nc;
  double qx[2], qy[2];
#pragma unroll
  for (int u = 0; u < 2; u++)
  {
    const int p = lane + 32 * u, c = p >> 2, g = p & 3;
    const double cxp = __shfl_sync(FULL, px, c & 31), cyp = __shfl_sync(FULL, py, c & 31);
    qx[u] = NB_ADD(cxp, NB_MUL((g & 2) ? 1.0 : -1.0, delta));
    qy[u] = NB_ADD(cyp, NB_MUL((g & 1) ? 1.0 : -1.0, delta));
    if (p < n4) pts[2 * p] = qx[u], pts[2 * p + 1] = qy[u];
  }
  __syncwarp();
  int rk[2] = { 0, 0 };
#pragma unroll 4
  for (int q = 0; q < n4; q++)
  {
    const double x = pts[2 * q], y = pts[2 * q + 1];
#pragma unroll
    for (int u = 0; u < 2; u++)
    {
      const int p = lane + 32 * u;
      const bool before = x < qx[u] || (x == qx[u] && (y < qy[u] || (y == qy[u] && q < p)));
      rk[u] += before ? 1 : 0;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; u++)
    if (lane + 32 * u < n4) srt[2 * rk[u]] = qx[u], srt[2 * rk[u] + 1] = qy[u];
  __syncwarp();
  // drop exact duplicates (they are adjacent)
  int m;
  {
    double sx[2], sy[2];
    bool keep[2];
#pragma unroll
    for (int u = 0; u < 2; u++)
    {
      const int p = lane + 32 * u;
      sx[u] = p < n4 ? srt[2 * p] : 0.0, sy[u] = p < n4 ? srt[2 * p + 1] : 0.0;
      keep[u] = p < n4 && (p == 0 || sx[u] != srt[2 * p - 2] || sy[u] != srt[2 * p - 1]);
    }
    const unsigned k0 = __ballot_sync(FULL, keep[0]), k1 = __ballot_sync(FULL, keep[1]);
    const unsigned below = (1u << lane) - 1u;
    const int pos0 = __popc(k0 & below), pos1 = __popc(k0) + __popc(k1 & below);
    m = __popc(k0) + __popc(k1);
    __syncwarp();
    if (keep[0]) srt[2 * pos0] = sx[0], srt[2 * pos0 + 1] = sy[0];
    if (keep[1]) srt[2 * pos1] = sx[1], srt[2 * pos1 + 1] = sy[1];
    __syncwarp();
  }
  if (m <= 2)
  {
    if (lane < m) hull[2 * lane] = srt[2 * lane], hull[2 * lane + 1] = srt[2 * lane + 1];
    return m;
  }
  int k = 0;
  if (lane < 2)
  {  // lane 0: lower chain over srt[0 .. m-1]; lane 1: upper chain over srt[m-1 .. 0]
    double* h = chn + lane * 2 * 66;
    for (int s = 0; s < m; s++)
    {
      const double* pt = srt + 2 * (lane ? m - 1 - s : s);
      const double x = pt[0], y = pt[1];
      while (k >= 2)
      {
        const double* o = h + 2 * (k - 2);
        const double* a = h + 2 * (k - 1);
        const double cr = NB_SUB(NB_MUL(NB_SUB(a[0], o[0]), NB_SUB(y, o[1])), NB_MUL(NB_SUB(a[1], o[1]), NB_SUB(x, o[0])));
        if (!(cr <= 0)) break;
        k--;
      }
      h[2 * k] = x, h[2 * k + 1] = y;
      k++;
    }
  }
  __syncwarp();
  const int kl = __shfl_sync(FULL, k, 0), ku = __shfl_sync(FULL, k, 1);
  const int cnt = kl + ku - 2;   // lower chain, then the upper one without its two end points
  for (int v = lane; v < cnt && v < NB_HMAX; v += 32)
  {
    const double* src = v < kl ? chn + 2 * v : chn + 2 * 66 + 2 * (v - kl + 1);
    hull[2 * v] = src[0], hull[2 * v + 1] = src[1];
  }
  return cnt;
}
#endif

// One interval of Neptune::SamplePointsOfIntervals (neptune.cpp:500-566): the S+1 samples of interval i when
// [t_start, t_end] is cut into num_pol intervals; out [S+1][2], idx [S+1] (optional)
// sample j of interval i; returns the piece it was taken from
NB_HD int nb_sample_one(const double* rec, double t_start, double t_end, int num_pol, int S, int i, int j, double* out)
{
  const int np = nb_rec_np(rec), nt = np + 1;
  const double* times = nb_rec_times(rec);
  const double* cx = nb_rec_coeff(rec, 0);
  const double* cy = nb_rec_coeff(rec, 1);
  const double deltaT = NB_SUB(t_end, t_start) / (1.0 * num_pol);
  const double ts = NB_ADD(NB_ADD(t_start, NB_MUL(deltaT, (double)i)), NB_MUL(deltaT / S, (double)j));
  const int low = nb_upper_bound(times, nt, ts);
  int ii;
  double te;
  if (low != nt)
  {
    ii = nb_sat(low - 1, 0, np - 1);
    te = NB_SUB(ts, times[ii]);
    if (te < 0)
      te = 0;
    else if (te > deltaT)
      te = deltaT;
  }
  else
  {
    const int k = low - 1;
    te = NB_SUB(times[k], times[k - 1]);
    ii = k - 1;
  }
  const double tv[4] = { NB_MUL(NB_MUL(te, te), te), NB_MUL(te, te), te, 1.0 };
  double x = 0, y = 0;
  for (int r = 0; r < 4; r++)
  {
    x = NB_ADD(x, NB_MUL(cx[4 * ii + r], tv[r]));
    y = NB_ADD(y, NB_MUL(cy[4 * ii + r], tv[r]));
  }
  out[0] = x;
  out[1] = y;
  return ii;
}

NB_HD void nb_sample_interval(const double* rec, double t_start, double t_end, int num_pol, int S, int i, double* out,
                              int* idx_out = nullptr)
{
  for (int j = 0; j <= S; j++)
  {
    const int ii = nb_sample_one(rec, t_start, t_end, num_pol, S, i, j, out + 2 * j);
    if (idx_out) idx_out[j] = ii;
  }
}

// Neptune::SamplePointsOfIntervals (neptune.cpp:500-566): out [num_pol][S+1][2], idx [num_pol][S+1]
NB_HD void nb_sample_points(const NbConsts& cs, const double* rec, double t_start, double t_end, double* out, int* idx_out)
{
  for (int i = 0; i < cs.num_pol; i++)
    nb_sample_interval(rec, t_start, t_end, cs.num_pol, cs.S, i, out + (size_t)i * (cs.S + 1) * 2,
                       idx_out ? idx_out + i * (cs.S + 1) : nullptr);
}

// ------------------------------------------------------------------------------------------ GJK
// argmax of d . v over the vertices, first maximum wins (gjk.cpp:30-47).  The dot products of four vertices are
// formed independently before the (sequential, order-preserving) comparisons, so that their latencies overlap.
NB_HD int nb_gjk_furthest(const double* v, int n, double dx, double dy)
{
  double best = NB_ADD(NB_MUL(dx, v[0]), NB_MUL(dy, v[1]));
  int idx = 0, i = 1;
  for (; i + 3 < n; i += 4)
  {
    const double pa = NB_ADD(NB_MUL(dx, v[2 * i]), NB_MUL(dy, v[2 * i + 1]));
    const double pb = NB_ADD(NB_MUL(dx, v[2 * i + 2]), NB_MUL(dy, v[2 * i + 3]));
    const double pc = NB_ADD(NB_MUL(dx, v[2 * i + 4]), NB_MUL(dy, v[2 * i + 5]));
    const double pd = NB_ADD(NB_MUL(dx, v[2 * i + 6]), NB_MUL(dy, v[2 * i + 7]));
    if (pa > best) best = pa, idx = i;
    if (pb > best) best = pb, idx = i + 1;
    if (pc > best) best = pc, idx = i + 2;
    if (pd > best) best = pd, idx = i + 3;
  }
  for (; i < n; i++)
  {
    const double p = NB_ADD(NB_MUL(dx, v[2 * i]), NB_MUL(dy, v[2 * i + 1]));
    if (p > best)
    {
      best = p;
      idx = i;
    }
  }
  return idx;
}
NB_HD void nb_gjk_support(const double* v1, int n1, const double* v2, int n2, double dx, double dy, double* s)
{
  const int i = nb_gjk_furthest(v1, n1, dx, dy), j = nb_gjk_furthest(v2, n2, -dx, -dy);
  s[0] = NB_SUB(v1[2 * i], v2[2 * j]);
  s[1] = NB_SUB(v1[2 * i + 1], v2[2 * j + 1]);
}
NB_HD double nb_dot2(const double* a, const double* b) { return NB_ADD(NB_MUL(a[0], b[0]), NB_MUL(a[1], b[1])); }
// b*(a.c) - a*(b.c)   (gjk.cpp:25-28)
NB_HD void nb_gjk_triple(const double* a, const double* b, const double* c, double* r)
{
  const double ac = nb_dot2(a, c), bc = nb_dot2(b, c);
  r[0] = NB_SUB(NB_MUL(b[0], ac), NB_MUL(a[0], bc));
  r[1] = NB_SUB(NB_MUL(b[1], ac), NB_MUL(a[1], bc));
}

// gjk::collision (gjk.cpp:76-148).  The simplex (at most three points, addressed by `index` in the reference)
// is kept in named registers s0, s1, s2 so that no thread needs a stack array.
NB_HD bool nb_gjk_collision(const double* v1, int n1, const double* v2, int n2)
{
  double s0[2], s1[2] = { 0, 0 }, s2[2], a[2], d[2], ao[2], ab[2], ac[2], abp[2], acp[2];
  double p1[2] = { 0, 0 }, p2[2] = { 0, 0 };
  int index = 0;
  for (int i = 0; i < n1; i++) p1[0] = NB_ADD(p1[0], v1[2 * i]), p1[1] = NB_ADD(p1[1], v1[2 * i + 1]);
  for (int i = 0; i < n2; i++) p2[0] = NB_ADD(p2[0], v2[2 * i]), p2[1] = NB_ADD(p2[1], v2[2 * i + 1]);
  d[0] = NB_SUB(p1[0] / n1, p2[0] / n2);
  d[1] = NB_SUB(p1[1] / n1, p2[1] / n2);
  if (d[0] == 0 && d[1] == 0) d[0] = 1.0;
  nb_gjk_support(v1, n1, v2, n2, d[0], d[1], s0);
  a[0] = s0[0], a[1] = s0[1];
  if (nb_dot2(a, d) <= 0) return false;
  d[0] = -a[0], d[1] = -a[1];
  for (int guard = 0; guard < 1000; guard++)
  {
    ++index;  // 1 or 2
    nb_gjk_support(v1, n1, v2, n2, d[0], d[1], a);
    if (index == 1)
      s1[0] = a[0], s1[1] = a[1];
    else
      s2[0] = a[0], s2[1] = a[1];
    if (nb_dot2(a, d) <= 0) return false;
    ao[0] = -a[0], ao[1] = -a[1];
    if (index < 2)
    {
      ab[0] = NB_SUB(s0[0], a[0]), ab[1] = NB_SUB(s0[1], a[1]);
      nb_gjk_triple(ab, ao, ab, d);
      if (sqrt(NB_ADD(NB_MUL(d[0], d[0]), NB_MUL(d[1], d[1]))) == 0) d[0] = ab[1], d[1] = -ab[0];
      continue;
    }
    ab[0] = NB_SUB(s1[0], a[0]), ab[1] = NB_SUB(s1[1], a[1]);
    ac[0] = NB_SUB(s0[0], a[0]), ac[1] = NB_SUB(s0[1], a[1]);
    nb_gjk_triple(ab, ac, ac, acp);
    if (nb_dot2(acp, ao) >= 0)
      d[0] = acp[0], d[1] = acp[1];
    else
    {
      nb_gjk_triple(ac, ab, ab, abp);
      if (nb_dot2(abp, ao) < 0) return true;
      s0[0] = s1[0], s0[1] = s1[1];
      d[0] = abp[0], d[1] = abp[1];
    }
    s1[0] = s2[0], s1[1] = s2[1];
    --index;
  }
  return false;
}

// One interval of Neptune::trajsAndPwpAreInCollision2d (neptune.cpp:781-802): my MINVO control points of
// interval i (P * A_rest_pos_basis_t_inverse_, :786-789) against the other agent's inflated hull over
// the same window.  Returns 1 collision, 0 free, -1 capacity.
NB_HD int nb_pwp_collides_interval(const NbConsts& cs, const double* coeff /*[3][8][4]*/, int n, int i, double t_start,
                                   const double* rec, double delta)
{
  const double t_end = NB_ADD(t_start, NB_MUL(cs.T, (double)n));
  const double deltaT = NB_SUB(t_end, t_start) / n;
  if (fabs(NB_SUB(deltaT, cs.T)) > 0.1) return 1;  // :772-780
  double A[8], hull[2 * NB_HMAX], nih0[2];
  int idx[2];
  for (int k = 0; k < 4; k++)
  {
    double x = 0, y = 0;
    for (int r = 0; r < 4; r++)
    {
      x = NB_ADD(x, NB_MUL(coeff[4 * i + r], cs.Ainv[r * 4 + k]));
      y = NB_ADD(y, NB_MUL(coeff[32 + 4 * i + r], cs.Ainv[r * 4 + k]));
    }
    A[2 * k] = x, A[2 * k + 1] = y;
  }
  const double w0 = NB_ADD(t_start, NB_MUL(deltaT, (double)i)), w1 = NB_ADD(t_start, NB_MUL(deltaT, (double)(i + 1)));
  const int hn = nb_hull_of_window(cs, rec, w0, w1, delta, hull, nih0, idx);
  if (hn < 0) return -1;
  return nb_gjk_collision(hull, hn < NB_HMAX ? hn : NB_HMAX, A, 4) ? 1 : 0;
}

// Neptune::trajsAndPwpAreInCollision2d (neptune.cpp:767-806): all intervals (used by the host emulation;
// the device kernel runs one interval per thread)
NB_HD int nb_pwp_collides(const NbConsts& cs, const double* coeff /*[3][8][4]*/, int n, double t_start, const double* rec,
                          double delta)
{
  for (int i = 0; i < n; i++)
  {
    const int r = nb_pwp_collides_interval(cs, coeff, n, i, t_start, rec, delta);
    if (r != 0) return r;
  }
  return 0;
}

// mu::composePieceWisePol (utils.cpp:318-402) on committed-trajectory records: the pieces of p1 that start
// after t and before p2 begins, then p2 (Neptune::replanFull neptune.cpp:1689-1699).  p1 / p2 are local
// copies whose first / last times are adjusted exactly as the reference adjusts its arguments (:320-336).
// Returns the number of pieces, 0 for the reference's empty "dummy" (:342-354), -1 beyond NB_TP pieces.
// nb_compose_records in two steps for a CTA: the (sequential, short) decision which pieces make up the result, by one
// thread, and the copying, by all of them.  plan: kind (0 all zero, 1 copy of p2 with times[0] = tm[0], 2 composed), the
// number of pieces, their knots tm[0 .. np] and for piece e its source (0: p1, 1: p2) and index there.  Returns what
// nb_compose_records returns (-1: more than NB_TP pieces).
struct NbComposePlan
{
  int kind, np;
  double tm[NB_TP + 1];
  unsigned char src[NB_TP], idx[NB_TP];
};

NB_HD int nb_compose_plan(double t, const double* p1_in, const double* p2_in, NbComposePlan* pl)
{
  const int n1 = (int)p1_in[0], n2 = (int)p2_in[0];
  pl->kind = 0, pl->np = 0;
  if (n1 < 1 || n2 < 1 || n1 > NB_TP || n2 > NB_TP) return 0;
  double t10 = p1_in[1], t1n = p1_in[1 + n1], t20 = p2_in[1];
  if (t > t1n && t < t20) t20 = t;
  if (t1n < t20) t20 = t1n;
  if (t < t10) t10 = t;
  if (fabs(t - t20) < 1e-5)
  {
    pl->kind = 1, pl->np = n2, pl->tm[0] = t20;
    return n2;
  }
  if (t1n < t20 || t > p2_in[1 + n2] || t < t10) return 0;
  int np = 0;
  pl->tm[0] = t;
  for (int i = 1; i <= n1; i++)  // i = 0 never qualifies: t1[0] <= t
  {
    const double ti = p1_in[1 + i];
    if (ti > t && ti < t20)
    {
      if (np >= NB_TP) return -1;
      pl->src[np] = 0, pl->idx[np] = (unsigned char)(i - 1);
      np++;
      pl->tm[np] = ti;
    }
  }
  for (int i = 0; i <= n2; i++)
  {
    const double ti = i == 0 ? t20 : p2_in[1 + i];
    if (ti > t)
    {
      if (np >= NB_TP) return -1;
      pl->src[np] = i == 0 ? 0 : 1, pl->idx[np] = (unsigned char)(i == 0 ? n1 - 1 : i - 1);
      np++;
      pl->tm[np] = ti;
    }
  }
  pl->kind = 2, pl->np = np;
  return np;
}

// element q of the composed record
NB_HD double nb_compose_fill(const NbComposePlan* pl, const double* p1_in, const double* p2_in, int q)
{
  if (pl->kind == 0) return 0.0;
  if (pl->kind == 1) return q == 1 ? pl->tm[0] : p2_in[q];
  if (q >= NB_REC_PWP) return p2_in[q];  // the message header is the new publication's
  if (q == 0) return (double)pl->np;
  if (q <= NB_TP + 1) return q - 1 <= pl->np ? pl->tm[q - 1] : 0.0;
  const int e = q - (NB_TP + 2), ax = e / (NB_TP * 4), rem = e - ax * NB_TP * 4, piece = rem >> 2, c = rem & 3;
  if (piece >= pl->np) return 0.0;
  const double* src = pl->src[piece] ? p2_in : p1_in;
  return src[1 + (NB_TP + 1) + ax * NB_TP * 4 + 4 * pl->idx[piece] + c];
}

NB_HD int nb_compose_records(double t, const double* p1_in, const double* p2_in, double* out)
{
  const int n1 = (int)p1_in[0], n2 = (int)p2_in[0];
  double t1[NB_TP + 1], t2[NB_TP + 1];
  if (n1 < 1 || n2 < 1 || n1 > NB_TP || n2 > NB_TP)  // empty pwp: the reference reads .back() of an empty vector
  {
    for (int q = 0; q < NB_REC; q++) out[q] = 0.0;
    return 0;
  }
  for (int i = 0; i <= NB_TP; i++) t1[i] = p1_in[1 + i], t2[i] = p2_in[1 + i];
  if (t > t1[n1] && t < t2[0]) t2[0] = t;
  if (t1[n1] < t2[0]) t2[0] = t1[n1];
  if (t < t1[0]) t1[0] = t;
  if (fabs(t - t2[0]) < 1e-5)
  {
    for (int q = 0; q < NB_REC; q++) out[q] = p2_in[q];
    out[1] = t2[0];
    return n2;
  }
  for (int q = 0; q < NB_REC; q++) out[q] = 0.0;
  if (t1[n1] < t2[0] || t > t2[n2] || t < t1[0]) return 0;
  int np = 0;
  out[1] = t;
  for (int i = 1; i <= n1; i++)  // i = 0 never qualifies: t1[0] <= t after :332-336
    if (t1[i] > t && t1[i] < t2[0])
    {
      if (np >= NB_TP) return -1;
      for (int ax = 0; ax < 3; ax++)
        for (int c = 0; c < 4; c++)
          out[1 + (NB_TP + 1) + ax * NB_TP * 4 + 4 * np + c] = p1_in[1 + (NB_TP + 1) + ax * NB_TP * 4 + 4 * (i - 1) + c];
      np++;
      out[1 + np] = t1[i];
    }
  for (int i = 0; i <= n2; i++)
    if (t2[i] > t)
    {
      if (np >= NB_TP) return -1;
      const double* src = (i == 0) ? p1_in : p2_in;
      const int k = (i == 0) ? n1 - 1 : i - 1;
      for (int ax = 0; ax < 3; ax++)
        for (int c = 0; c < 4; c++)
          out[1 + (NB_TP + 1) + ax * NB_TP * 4 + 4 * np + c] = src[1 + (NB_TP + 1) + ax * NB_TP * 4 + 4 * k + c];
      np++;
      out[1 + np] = t2[i];
    }
  out[0] = (double)np;
  for (int q = NB_REC_PWP; q < NB_REC; q++) out[q] = p2_in[q];  // the message header is the new publication's
  return np;
}
