// nb_prune.cuh -- exact removal of redundant separating lines of one (agent, interval).
//
// Every accepted line l of interval i constrains the SAME four control points q_k:
//     n_l . q_k <= c_l  (c_l = 1 - d_l)          [solver_gurobi_poly.cpp:485-489 and siblings]
// so per interval the lines describe one convex polygon and only its facets matter.  The reference
// hands all rows to Gurobi (whose presolve drops dominated rows); here the redundant half-planes are
// removed geometrically, which leaves the feasible set -- and therefore the unique QP optimum --
// unchanged.  Polar duality about a strictly interior point p0 (the centroid of the initial control
// points: every line has n.q + d <= -1 there): line l maps to the dual point n_l / (c_l - n_l.p0), and
// l is redundant iff its dual point lies inside conv(all dual points + origin).  The hull is found by
// gift wrapping; only points STRICTLY inside it (by a tolerance) are dropped, so rounding can only
// keep a redundant line, never drop a needed one.
#pragma once
#include "nb_common.cuh"

#define NB_PRUNE_KMAX 48  // more hull vertices than this: keep every line

struct NbPruneShared
{
  double* px;     // [LS+1] dual points (slot-indexed), index LS = origin
  double* py;
  uint8_t* valid; // [LS+1]
  int* red;       // [NT] per-thread candidates
  int* hull;      // [NB_PRUNE_KMAX+1]
  int* misc;      // [8]
};

template <int NT>
struct Cta
{
  int tid;
  NB_HD Cta(int t) : tid(t) {}
#if defined(__CUDA_ARCH__)
  NB_DEV void sync() const { __syncthreads(); }
#else
  void sync() const {}
#endif
};

// a beats b as the next CCW hull vertex after c: b is to the left of c->a, or collinear and nearer
NB_HD bool nb_wrap_better(const double* px, const double* py, int c, int a, int b)
{
  if (b < 0) return true;
  if (a < 0) return false;
  const double ax = px[a] - px[c], ay = py[a] - py[c], bx = px[b] - px[c], by = py[b] - py[c];
  const double cr = ax * by - ay * bx;  // > 0: b is left of c->a  => a is the more clockwise => a wins
  if (cr > 0) return true;
  if (cr < 0) return false;
  const double da = ax * ax + ay * ay, db = bx * bx + by * by;
  if (da != db) return da > db;  // collinear: farther wins
  return a < b;
}

template <int NT>
NB_HD int nb_cta_reduce_wrap(const Cta<NT>& cta, const NbPruneShared& ps, int c, int mine)
{
  constexpr int NW = (NT + 31) / 32;
#if defined(__CUDA_ARCH__)
  // butterfly inside each warp (nb_wrap_better is a strict total order, so every lane converges on the
  // same winner), then one shared-memory exchange between the warps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const int other = __shfl_xor_sync(0xffffffffu, mine, o);
    if (other != mine && nb_wrap_better(ps.px, ps.py, c, other, mine)) mine = other;
  }
  if ((cta.tid & 31) == 0) ps.misc[cta.tid >> 5] = mine;
  cta.sync();
#else
  ps.misc[0] = mine;
#endif
  int best = -1;
  for (int w = 0; w < NW; w++)
    if (ps.misc[w] != best && nb_wrap_better(ps.px, ps.py, c, ps.misc[w], best)) best = ps.misc[w];
  cta.sync();
  return best;
}

// lexicographic "a before b" on dual points, ties by slot index
NB_HD bool nb_lex_before(const double* px, const double* py, int a, int b)
{
  if (b < 0) return true;
  if (a < 0) return false;
  if (px[a] != px[b]) return px[a] < px[b];
  if (py[a] != py[b]) return py[a] < py[b];
  return a < b;
}

// ok[s] == 1 marks a solved line in slot s; keep[s] is set to 1 for the lines the QP must see.
template <int NT>
NB_HD void nb_prune_lines(const Cta<NT>& cta, int LS, const double* lines, const uint8_t* ok, const double cp[8],
                          const NbPruneShared& ps, uint8_t* keep)
{
  const double p0x = 0.25 * (cp[0] + cp[2] + cp[4] + cp[6]), p0y = 0.25 * (cp[1] + cp[3] + cp[5] + cp[7]);
  int cnt = 0;
  double smax = 0.0;
  for (int s = cta.tid; s <= LS; s += NT)
  {
    uint8_t v = 0;
    double x = 0.0, y = 0.0;
    if (s == LS)
      v = 1;  // the origin: the polygon of lines alone may be unbounded
    else if (ok[s] == 1)
    {
      const double n0 = lines[3 * s], n1 = lines[3 * s + 1], c = 1.0 - lines[3 * s + 2];
      const double r = c - n0 * p0x - n1 * p0y;
      if (r > 0.0)
      {
        v = 1;
        x = n0 / r;
        y = n1 / r;
        cnt++;
        smax = fmax(smax, fmax(fabs(x), fabs(y)));
      }
      else
        v = 2;  // cannot happen for a solved line (r >= 2); keep it untouched
    }
    ps.px[s] = x;
    ps.py[s] = y;
    ps.valid[s] = v;
    if (s < LS) keep[s] = (v != 0) ? 1 : 0;
  }
  cta.sync();
  // lexicographically smallest valid point and the number of valid lines
  int mine = -1, cnt_local = 0;
  for (int s = cta.tid; s <= LS; s += NT)
    if (ps.valid[s] == 1)
    {
      if (s < LS) cnt_local++;
      if (nb_lex_before(ps.px, ps.py, s, mine)) mine = s;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    const int other = __shfl_xor_sync(0xffffffffu, mine, o);
    cnt_local += __shfl_xor_sync(0xffffffffu, cnt_local, o);
    if (other != mine && nb_lex_before(ps.px, ps.py, other, mine)) mine = other;
  }
  if ((cta.tid & 31) == 0)
  {
    ps.misc[cta.tid >> 5] = mine;
    ps.red[cta.tid >> 5] = cnt_local;
  }
  cta.sync();
#else
  ps.misc[0] = mine;
  ps.red[0] = cnt_local;
#endif
  int start = -1, total = 0;
  for (int w = 0; w < (NT + 31) / 32; w++)
  {
    total += ps.red[w];
    if (ps.misc[w] != start && nb_lex_before(ps.px, ps.py, ps.misc[w], start)) start = ps.misc[w];
  }
  cta.sync();
  if (cta.tid == 0) ps.hull[0] = start;
  cta.sync();
  if (total <= 3) return;  // nothing worth pruning
  // gift wrapping, counter-clockwise
  int k = 1, cur = start;
  bool overflow = false;
  while (true)
  {
    int cand = -1;
    for (int s = cta.tid; s <= LS; s += NT)
      if (ps.valid[s] == 1 && s != cur && (ps.px[s] != ps.px[cur] || ps.py[s] != ps.py[cur]) &&
          nb_wrap_better(ps.px, ps.py, cur, s, cand))
        cand = s;
    const int nxt = nb_cta_reduce_wrap<NT>(cta, ps, cur, cand);
    if (nxt < 0 || nxt == start || (ps.px[nxt] == ps.px[start] && ps.py[nxt] == ps.py[start])) break;
    if (k >= NB_PRUNE_KMAX)
    {
      overflow = true;
      break;
    }
    if (cta.tid == 0) ps.hull[k] = nxt;
    k++;
    cur = nxt;
    cta.sync();
  }
  cta.sync();
  if (overflow || k < 3) return;
  // drop the lines whose dual point is strictly inside the hull
  double scale = 0.0;
  for (int q = 0; q < k; q++) scale = fmax(scale, fmax(fabs(ps.px[ps.hull[q]]), fabs(ps.py[ps.hull[q]])));
  const double tol = 1e-9 * scale * scale;
  for (int s = cta.tid; s < LS; s += NT)
  {
    if (ps.valid[s] != 1) continue;
    bool inside = true;
    for (int q = 0; q < k && inside; q++)
    {
      const int a = ps.hull[q], b = ps.hull[q + 1 < k ? q + 1 : 0];
      const double cr = (ps.px[b] - ps.px[a]) * (ps.py[s] - ps.py[a]) - (ps.py[b] - ps.py[a]) * (ps.px[s] - ps.px[a]);
      if (!(cr > tol)) inside = false;
    }
    if (inside) keep[s] = 0;
  }
  (void)cnt;
  (void)smax;
}
