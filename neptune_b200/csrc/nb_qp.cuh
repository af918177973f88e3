// nb_qp.cuh -- K4: batched trajectory QP, ONE AGENT PER WARP.
//
// Replaces PolySolverGurobi::optimize's two m_.optimize() calls and the model they are run on
// (reference neptune/src/solver_gurobi_poly.cpp:322-383 addObjective, :385-471 + :659-708
// addConstraints, :804-887 optimize).  The QP is solved in the reduced
// coordinates w of NbQpTable (equalities eliminated on the host once) by an infeasible-start
// Mehrotra predictor-corrector interior-point method.  Every inequality row is a function of one
// "feature" (a MINVO position / velocity control point or an end acceleration of one axis and
// interval: bound rows :441-470) or of the (x,y) pair of one position control point (separating-line
// rows :485-489, :546-550, :587-591, :754-758), so the normal matrix is assembled as
// Hr + C^T W C with W diagonal plus one 2x2 coupling per control point -- no per-row outer products.
//
// Execution model: one CTA of four warps per agent.  The data-parallel phases (row sweeps, feature products, the
// assembly of the normal matrix) are lane-strided over the 128 threads with at most two rows per thread; the inherently
// sequential part -- factorisation and triangular solves -- runs on warp 0 alone with the matrix rows in REGISTERS
// and pivots travelling by shuffle, so it costs no block barrier.  The normal matrix is block diagonal, an (x,y) block
// of 2 dof <= 16 rows and a z block of dof <= 8 rows (the separating lines couple x with y only); the two blocks
// are factorised (L D L^T) and solved IN LOCKSTEP by disjoint lanes, which shortens the sequential pivot chain
// from 3 dof to 2 dof columns.  The quadratic terminal row (:697-702), which couples all three axes through a
// rank-one term, enters by the Sherman-Morrison formula.  The corrector's right-hand side is obtained from two
// per-feature sums gathered during the predictor sweep, so an iteration makes three sweeps over the rows
// (residual, predictor direction, final direction).
//
// Lane-strided SPMD phases; compiles with NL = 128 (one CTA) on the device and NL = 1 in the host emulation.
#pragma once
#include "nb_common.cuh"

#define NB_KLD 25                      // row stride of the normal matrix in shared memory (odd: conflict-free rows)
#define NB_NF3 (3 * NB_NFEAT_AX)       // features of the three axes
#define NB_ROW_LINE0 (2 * NB_NF3)      // first separating-line row: upper bound rows [0, 192), lower [192, 384)
#define NB_NELEM (2 * NB_DOF_MAX * (2 * NB_DOF_MAX - 1) / 2 + NB_DOF_MAX * (NB_DOF_MAX - 1) / 2)  // 120 + 28

struct NbQpShared
{
  double y[NB_NF3];    // feature values
  double dy[NB_NF3];   // feature values of a direction
  double du[NB_NF3];   // dual load per feature (G^T lambda reduced to features), for r_d
  double La[NB_NF3];   // primal load per feature of the predictor right-hand side
  double A1[NB_NF3];   // corrector: sum of ds_aff dl_aff / s per feature
  double A2[NB_NF3];   // corrector: sum of 1 / s per feature
  double omv[NB_NFEAT_AX * 4];  // weights per feature [f][0..2] = axis x, y, z; [f][3] = x-y coupling (position CPs)
  double K[NB_NV_MAX * NB_KLD];
  double col[NB_NV_MAX], invd[NB_NV_MAX];
  double w[NB_NV_MAX], dw[NB_NV_MAX], rd[NB_NV_MAX], g0[NB_NV_MAX], gq[NB_NV_MAX], gobj[NB_NV_MAX], uq[NB_NV_MAX];
  double init3[3][3], pf[3];
  double xin[3][4 * NB_NPOL];
  double blo[24], bhi[24];                          // bounds of feature kind (axis, j)
  unsigned char ax_of[NB_NV_MAX], c_of[NB_NV_MAX];  // variable a -> (axis, column)
  unsigned char pca[NB_NPAIR], pcb[NB_NPAIR];       // pair p -> (ca, cb), ca >= cb
  unsigned char ei[NB_NELEM], ej[NB_NELEM];         // strictly-lower elements (i > j) of the two diagonal blocks
  int nelem;
};

// Reciprocal for the interior-point algebra (pivots, 1 / s): hardware seed (MUFU.RCP64H, ~20 bits) and two Newton
// steps -- full double precision to within an ulp or two, a third of the instructions of the correctly rounded
// __drcp_rn and no slow-path branch.  Arguments are never zero, denormal or infinite here.
NB_HD double nb_rcp(double x)
{
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

struct NbQpRows  // per-agent row state (shared memory when it fits, else global scratch)
{
  double* s;    // [384 + 4*LCAP]
  double* lam;
  double* dsa;
  double* dla;
  double* inv;  // 1/s of the current iterate (written by the RESID pass)
  const double* const* clb;    // [8] kept lines (n0, n1, c = 1 - d): line l of interval i is clb[i] + 3 l, for
                               // lstart[i] <= l < lstart[i+1] (one contiguous list, or one run per interval)
  const int* lstart;           // [n+1] first line of each interval
  const unsigned char* l2i;    // optional [L]: interval of line l (shared-memory path)
};

NB_HD void nb_feat_bounds(const NbConsts& cs, int ax, int j, double& lo, double& hi)
{
  if (j < 4)
  {
    lo = cs.lim_min[ax];
    hi = cs.lim_max[ax];
  }
  else if (j < 7)
  {
    lo = -cs.v_max;
    hi = cs.v_max;
  }
  else
  {
    lo = -cs.a_max;
    hi = cs.a_max;
  }
}

// y = c0 . init3 + C w  (with_const) or dy = C dw.  Lanes over features: C is read transposed.
template <int NL>
NB_HD void nb_qp_features(const Group<NL>& g, const NbQpTable* tb, const NbQpShared* sh_c, double* out,
                          const double* vec, bool with_const)
{
  const int n = tb->n, dof = tb->dof;
  for (int q = g.lane; q < NB_NF3; q += NL)
  {
    const int ax = q >> 6, fl = q & 63;
    if (fl >= 8 * n) continue;
    double v = 0.0;
    if (with_const)
      v = tb->c0[fl][0] * sh_c->init3[ax][0] + tb->c0[fl][1] * sh_c->init3[ax][1] + tb->c0[fl][2] * sh_c->init3[ax][2];
#pragma unroll
    for (int c = 0; c < NB_DOF_MAX; c++)
      if (c < dof) v += tb->Ct[c][fl] * vec[ax * dof + c];
    out[q] = v;
  }
}

enum
{
  NB_PASS_RESID = 0,  // residuals, weights, predictor load; sums mu, |rp|max
  NB_PASS_DIR_PRED,   // predictor direction: ds, dl -> scratch; step length, mu_aff sums; corrector sums A1
  NB_PASS_DIR_CORR,   // final direction: ds, dl -> scratch; step length
  NB_PASS_START       // Nocedal-Wright start: s = max(1,|s+ds|), lam likewise
};

struct NbPassAcc
{
  double sum_sl, rp_max, rmax, sum_cross, sum_dd;  // rmax = max over rows of (-ds/s, -dl/lam): alpha_max = 1/rmax
};

// One inequality row r: v = row value (<= 0 wanted), gd = row . direction.
//   RESID   : applies the pending step, returns the predictor load tau = (lam rp - s lam) / s, wgt = lam / s, iv = 1 / s
//   DIR_PRED: returns ds_aff dl_aff / s (the corrector's second-order term)
template <int MODE>
NB_HD double nb_row(const NbQpRows& R, int r, double v, double gd, double sigmu, double alpha, NbPassAcc& acc,
                    double& wgt, double& iv)
{
  double s = R.s[r], lam = R.lam[r];
  wgt = 0.0;
  iv = 0.0;
  if (MODE == NB_PASS_RESID)
  {
    if (alpha != 0.0)
    {  // pending step of the previous iteration (s, lam) += alpha (ds, dl); v already is at the new point
      s += alpha * R.dsa[r];
      lam += alpha * R.dla[r];
      R.s[r] = s;
      R.lam[r] = lam;
    }
    const double rp = v + s;
    const double inv = nb_rcp(s);
    R.inv[r] = inv;
    iv = inv;
    wgt = lam * inv;
    acc.sum_sl += s * lam;
    acc.rp_max = fmax(acc.rp_max, fabs(rp));
    return (lam * rp - s * lam) * inv;  // predictor: rc = s lam
  }
  if (MODE == NB_PASS_DIR_PRED)
  {
    const double inv = R.inv[r];
    const double ds = -(v + s) - gd;
    const double dl = (-s * lam - lam * ds) * inv;
    acc.rmax = fmax(acc.rmax, -ds * inv);       // -ds/s
    acc.rmax = fmax(acc.rmax, 1.0 + ds * inv);  // -dl/lam = (s + ds)/s for the predictor
    acc.sum_cross += s * dl + lam * ds;
    acc.sum_dd += ds * dl;
    R.dsa[r] = ds;
    R.dla[r] = dl;
    return ds * dl * inv;
  }
  if (MODE == NB_PASS_DIR_CORR)
  {
    const double inv = R.inv[r];
    const double rc = s * lam + R.dsa[r] * R.dla[r] - sigmu;
    const double ds = -(v + s) - gd;
    const double dl = (-rc - lam * ds) * inv;
    acc.rmax = fmax(acc.rmax, -ds * inv);
    acc.rmax = fmax(acc.rmax, -dl * nb_rcp(lam));
    R.dsa[r] = ds;
    R.dla[r] = dl;
    return 0.0;
  }
  // NB_PASS_START
  {
    const double a = fabs(s + R.dsa[r]), b = fabs(lam + R.dla[r]);
    R.s[r] = a > 1.0 ? a : 1.0;
    R.lam[r] = b > 1.0 ? b : 1.0;
  }
  return 0.0;
}

// One sweep over every inequality row.  Bound rows are visited feature by feature (both sides of a
// feature by the same lane), line rows control point by control point (item = interval i, control point k).
// All per-feature accumulations are conflict-free and in a fixed order.
template <int NL, int MODE>
NB_HD void nb_qp_pass(const Group<NL>& g, const NbQpTable* tb, NbQpShared* sh, const NbQpRows& R, int nlines, double sigmu,
                      double alpha, NbPassAcc& acc)
{
  const int n = tb->n;
  for (int q = g.lane; q < NB_NF3; q += NL)
  {
    const int ax = q >> 6, fl = q & 63, f = q;
    if (fl >= 8 * n) continue;
    double w_u, w_l, i_u, i_l;
    const double lo = sh->blo[ax * 8 + (fl & 7)], hi = sh->bhi[ax * 8 + (fl & 7)];
    double yv = sh->y[f];
    const double dv = sh->dy[f];
    if (MODE == NB_PASS_RESID && alpha != 0.0)
    {  // features are linear in w: y(w + alpha dw) = y + alpha dy
      yv += alpha * dv;
      sh->y[f] = yv;
    }
    const double t_u = nb_row<MODE>(R, f, yv - hi, dv, sigmu, alpha, acc, w_u, i_u);
    const double t_l = nb_row<MODE>(R, NB_NF3 + f, lo - yv, -dv, sigmu, alpha, acc, w_l, i_l);
    if (MODE == NB_PASS_RESID)
    {
      sh->La[f] = t_u - t_l;
      sh->A2[f] = i_u - i_l;
      sh->omv[fl * 4 + ax] = w_u + w_l;
      sh->du[f] = R.lam[f] - R.lam[NB_NF3 + f];
    }
    if (MODE == NB_PASS_DIR_PRED) sh->A1[f] = t_u - t_l;
  }
  g.sync();
  // separating-line rows, one row per lane and step: row (l, k) is line l on control point k of its interval
  for (int q = g.lane; q < 4 * nlines; q += NL)
  {
    const int l = q >> 2, k = q & 3;
    int i;
    if (R.l2i)
      i = R.l2i[l];
    else
    {
      i = 0;
      for (int j = 1; j < n; j++) i += (l >= R.lstart[j]) ? 1 : 0;
    }
    const double* cl = R.clb[i] + 3 * l;
    const double n0 = cl[0], n1 = cl[1], c = cl[2];
    const int fl = i * 8 + k;
    const double yx = sh->y[fl], yy = sh->y[NB_NFEAT_AX + fl];
    const double dx = (MODE == NB_PASS_RESID) ? 0.0 : sh->dy[fl], dyv = (MODE == NB_PASS_RESID) ? 0.0 : sh->dy[NB_NFEAT_AX + fl];
    const int r = NB_ROW_LINE0 + q;
    double wgt, iv;
    const double t = nb_row<MODE>(R, r, n0 * yx + n1 * yy - c, n0 * dx + n1 * dyv, sigmu, alpha, acc, wgt, iv);
    if (MODE == NB_PASS_RESID)
    {  // parked for the sums below (the pending step has consumed dsa / dla of this row already)
      R.dsa[r] = t;
      R.dla[r] = wgt;
    }
  }
  // this lane's share of the sweep's sums is complete: one partial per warp, folded by the caller after the next barrier
  if (MODE == NB_PASS_RESID) g.put(0, acc.sum_sl, acc.rp_max, 0.0);
  if (MODE == NB_PASS_DIR_PRED) g.put(1, acc.sum_cross, acc.rmax, acc.sum_dd);
  if (MODE == NB_PASS_DIR_CORR) g.put(0, 0.0, acc.rmax, 0.0);
  if (MODE == NB_PASS_START) return;
  g.sync();
  if (MODE == NB_PASS_DIR_CORR) return;
  // per control point (i, k): what its line rows load on the x and y features.  RESID: four lanes per control point, one
  // per kind of sum (predictor load, 1 / s, weights, multipliers); PRED: the lanes of a control point split its lines
  constexpr int SUB = Group<NL>::SUB;
  if (MODE == NB_PASS_RESID)
  {
    for (int q = g.lane; q < 16 * n; q += NL)
    {
      const int cp = q >> 2, role = q & 3, i = cp >> 2, k = cp & 3;
      const int fl = i * 8 + k, fx = fl, fy = NB_NFEAT_AX + fl;
      const double* cl = R.clb[i];
      const double* src = role == 0 ? R.dsa : (role == 1 ? R.inv : (role == 2 ? R.dla : R.lam));
      double s0 = 0, s1 = 0, s2 = 0;
      const int l1 = R.lstart[i + 1];
      for (int l = R.lstart[i]; l < l1; l += 4)
      {  // four lines per trip, loads first (same order of the sums as one line per trip)
        double n0[4], n1[4], v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
          const bool in = l + u < l1;
          n0[u] = in ? cl[3 * (l + u)] : 0.0, n1[u] = in ? cl[3 * (l + u) + 1] : 0.0;
          v[u] = in ? src[NB_ROW_LINE0 + 4 * (l + u) + k] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
          const double a = role == 2 ? v[u] * n0[u] : v[u];
          s0 += a * n0[u];
          s1 += a * n1[u];
          s2 += v[u] * n1[u] * n1[u];
        }
      }
      if (role == 0)
        sh->La[fx] += s0, sh->La[fy] += s1;
      else if (role == 1)
        sh->A2[fx] += s0, sh->A2[fy] += s1;
      else if (role == 2)
        sh->omv[fl * 4 + 0] += s0, sh->omv[fl * 4 + 1] += s2, sh->omv[fl * 4 + 3] = s1;
      else
        sh->du[fx] += s0, sh->du[fy] += s1;
    }
  }
  else
  {
    const int items = 4 * n * SUB;  // <= NL for SUB > 1, so every lane reaches the shuffles below
    for (int b0 = 0; b0 < items; b0 += NL)
    {
      const int q = b0 + g.lane;
      const bool on = q < items;
      const int cp = on ? q / SUB : 0, sub = q % SUB, i = cp >> 2, k = cp & 3;
      const int fl = i * 8 + k;
      const double* cl = R.clb[i];
      double lx = 0, ly = 0;
      if (on)
        for (int l = R.lstart[i] + sub; l < R.lstart[i + 1]; l += SUB)
        {
          const int r = NB_ROW_LINE0 + 4 * l + k;
          const double t = R.dsa[r] * R.dla[r] * R.inv[r];
          lx += t * cl[3 * l];
          ly += t * cl[3 * l + 1];
        }
      lx = g.sub_sum(lx), ly = g.sub_sum(ly);
      if (on && sub == 0) sh->A1[fl] += lx, sh->A1[NB_NFEAT_AX + fl] += ly;
    }
  }
  g.sync();
}

// dst[a] = bsign * base[a] + sign * (C^T vecF)[a] + es * extra[a]   (a = ax*dof + c); item (a, sub-lane): the SUB
// adjacent lanes of a variable split its features and combine by shuffle
template <int NL>
#if defined(__CUDA_ARCH__)
__device__ __noinline__   // one copy for the three calls of an iteration
#else
inline
#endif
void nb_qp_ct_all(const Group<NL>& g, const NbQpTable* tb, const NbQpShared* sh_t, const double* vecF, double* dst,
                  const double* base, double bsign, double sign, const double* extra, double es)
{
  constexpr int SUB = Group<NL>::SUB;
  const int n = tb->n, dof = tb->dof, nv = 3 * dof;
  const int items = nv * SUB;  // <= NL for SUB > 1, so every lane reaches the shuffles below
  for (int b0 = 0; b0 < items; b0 += NL)
  {
    const int q = b0 + g.lane;
    const bool on = q < items;
    const int a = on ? q / SUB : 0, sub = q % SUB;
    const int ax = (a >= dof ? 1 : 0) + (a >= 2 * dof ? 1 : 0), c = a - ax * dof;
    const double tail = bsign * base[a] + (extra ? extra[a] * es : 0.0);   // loaded before the sums, not after them
    double v0 = 0.0, v1 = 0.0;
    if (on)
    {
      const double* vf = vecF + ax * NB_NFEAT_AX + sub;
      const double* cc = &tb->C[sub][c];
      const int cnt = (8 * n - sub + SUB - 1) / SUB;   // features sub, sub + SUB, ...
      // fully unrolled with predicates: every load is issued before the first product (a shared-memory load costs ~50
      // cycles when it sits in the dependent chain, and a rolled loop with a run-time bound leaves it there)
      constexpr int MAXC = NB_NFEAT_AX / SUB;
      double cv[MAXC], fv[MAXC];
#pragma unroll
      for (int k = 0; k < MAXC; k++)
      {
        const bool in = k < cnt;
        cv[k] = in ? cc[(size_t)k * SUB * NB_DOF_MAX] : 0.0;
        fv[k] = in ? vf[k * SUB] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < MAXC; k += 2)
      {
        v0 += cv[k] * fv[k];
        if (k + 1 < MAXC) v1 += cv[k + 1] * fv[k + 1];
      }
    }
    const double v = g.sub_sum(v0 + v1);
    if (on && sub == 0) dst[a] = tail + sign * v;
  }
}

// Normal matrix K0 = Hr + C^T W C + lam_q Hess(c) (lower triangles of the (x,y) block and of the z block; the
// rank-one term of the quadratic row is NOT included, see nb_qp_solve_blocks).  Items = (block, pair): blocks 0..2 are
// the same-axis blocks, block 3 the y-x coupling (symmetric: both axes share C), all through the products PP.
template <int NL>
NB_HD void nb_qp_assemble(const Group<NL>& g, const NbQpTable* tb, NbQpShared* sh, bool has_qc, double lam_q)
{
  const int n = tb->n, dof = tb->dof, np = dof * (dof + 1) / 2;
  constexpr int S2 = NL >= 128 ? 2 : 1;  // adjacent lanes that split the features of one (block, pair)
  const int items = 4 * np * S2;
  for (int b0 = 0; b0 < items; b0 += NL)
  {
    const int q = b0 + g.lane;
    const bool on = q < items;
    const int e = on ? q / S2 : 0, sub = q % S2;
    const int blk = e / np, p = e - blk * np;
    const int ca = sh->pca[p], cb = sh->pcb[p];
    double v0 = 0.0, v1 = 0.0;
    if (on)
    {
      const double* om = sh->omv + blk;
#pragma unroll 2
      for (int fl = 2 * sub; fl < 8 * n; fl += 2 * S2)
      {
        v0 += om[fl * 4] * tb->PP[fl][p];
        v1 += om[fl * 4 + 4] * tb->PP[fl + 1][p];
      }
    }
    double v = v0 + v1;
#if defined(__CUDA_ARCH__)
    if (S2 == 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
#endif
    if (!on || sub != 0) continue;
    if (blk < 3)
    {
      v += tb->Hr[ca][cb];
      if (has_qc) v += lam_q * 2.0 * tb->tq[ca] * tb->tq[cb];
      sh->K[(blk * dof + ca) * NB_KLD + blk * dof + cb] = v;
    }
    else
    {
      sh->K[(dof + ca) * NB_KLD + cb] = v;
      sh->K[(dof + cb) * NB_KLD + ca] = v;
    }
  }
  g.sync();
}

// Factorisation K = L D L^T of the two diagonal blocks [0, nxy) and [nxy, nxy + nz) in lockstep, element-parallel:
// lane e owns the strictly-lower element (i, j); at step t the elements right of column t (of their block) take the
// rank-one update with the UNSCALED column, K[i][j] -= K[i][k] K[j][k] / K[k][k], so column k and the pivot are only
// read during the step and one barrier per column suffices.  Diagonal entries are updated by the lanes j == i of a
// second list range.  Afterwards K[i][k] (i > k) = L[i][k] and invd[k] = 1 / D[k].
template <int NL>
NB_HD void nb_qp_factor_blocks(const Group<NL>& g, NbQpShared* sh, int nxy, int nz)
{
  double* K = sh->K;
  const int nv = nxy + nz, ne = sh->nelem;
  if (NL >= 128)
  {  // a thread owns at most two elements (ne + nv <= 172): their addresses and last step are fixed over the columns
    double *pe[2], *pi[2], *pj[2], *pd[2];
    int last[2];  // the element takes the updates of steps t < last
#pragma unroll
    for (int u = 0; u < 2; u++)
    {
      const int e = g.lane + u * NL;
      const bool on = e < ne + nv;
      const int i = !on ? 0 : (e < ne ? sh->ei[e] : e - ne), j = !on ? 0 : (e < ne ? sh->ej[e] : e - ne);
      const int base = i >= nxy ? nxy : 0;
      pe[u] = K + i * NB_KLD + j, pi[u] = K + i * NB_KLD + base, pj[u] = K + j * NB_KLD + base, pd[u] = K + base * (NB_KLD + 1);
      last[u] = on ? j - base : 0;
    }
#pragma unroll 1
    for (int t = 0; t < nxy; t++)
    {
#pragma unroll
      for (int u = 0; u < 2; u++)
        if (t < last[u])
        {
          double d = pd[u][t * (NB_KLD + 1)];  // K[k][k], k = base + t
          d = d > 1e-300 ? d : 1e-300;  // rank-deficiency guard (K is positive definite by construction)
          *pe[u] -= pi[u][t] * pj[u][t] * nb_rcp(d);
        }
      g.sync();
    }
  }
  else
  {
#pragma unroll 1
    for (int t = 0; t < nxy; t++)
    {
#pragma unroll 1
      for (int e = g.lane; e < ne + nv; e += NL)
      {
        const int i = e < ne ? sh->ei[e] : e - ne, j = e < ne ? sh->ej[e] : e - ne;
        const bool second = i >= nxy;
        const int k = second ? nxy + t : t;
        if ((second && t >= nz) || j <= k) continue;
        double d = K[k * NB_KLD + k];
        d = d > 1e-300 ? d : 1e-300;  // rank-deficiency guard (K is positive definite by construction)
        K[i * NB_KLD + j] -= K[i * NB_KLD + k] * K[j * NB_KLD + k] * nb_rcp(d);
      }
      g.sync();
    }
  }
  for (int k = g.lane; k < nv; k += NL)
  {
    double d = K[k * NB_KLD + k];
    d = d > 1e-300 ? d : 1e-300;
    sh->invd[k] = nb_rcp(d);
  }
  g.sync();
  for (int e = g.lane; e < ne; e += NL) K[sh->ei[e] * NB_KLD + sh->ej[e]] *= sh->invd[sh->ej[e]];
  g.sync();
}

#if defined(__CUDA_ARCH__)
// x_i of K0 x = b (lane i of warp 0), both blocks in lockstep: 2 nxy dependent shuffle + FMA steps in two short loops
// (they stay in the instruction cache: a CTA of four warps runs this code once per solve, and straight-line code that
// has fallen out of the cache is fetched from L2); the entry of L a step needs is loaded while the pivot travels.
__device__ __noinline__ double nb_warp_solve(const double* K, const double* invd, int nxy, int nz, int lane, double x)
{
  const bool second = lane >= nxy;
  const int base = second ? nxy : 0, nb = second ? nz : nxy, il = lane - base;
  const bool mine = il < nb;
  const double* row = K + lane * NB_KLD + base;   // L[i][base + t], t < il
  const double* col = K + base * NB_KLD + lane;   // L[base + t][i], t > il (stride NB_KLD)
  // blocks of four steps; the entries of L of the NEXT block are loaded while this one runs, so that the chain is
  // shuffle + FMA only (steps past the block's size carry l = 0 and change nothing)
  const int nblk = (nxy + 3) >> 2;
  const int tf = mine ? il : 0, tb0 = mine ? il : 1 << 20, tb1 = mine ? nb : 0;   // forward: t < tf; backward: tb0 < t < tb1
  double l0 = 0 < tf ? row[0] : 0.0, l1 = 1 < tf ? row[1] : 0.0, l2 = 2 < tf ? row[2] : 0.0, l3 = 3 < tf ? row[3] : 0.0;
  const double id = mine ? invd[lane] : 0.0;
#pragma unroll 1
  for (int k = 0; k < nblk; k++)
  {
    const int t = 4 * k, u = t + 4;
    const double m0 = u < tf ? row[u] : 0.0, m1 = u + 1 < tf ? row[u + 1] : 0.0, m2 = u + 2 < tf ? row[u + 2] : 0.0,
                 m3 = u + 3 < tf ? row[u + 3] : 0.0;
    x -= l0 * __shfl_sync(0xffffffffu, x, (base + t) & 31);
    x -= l1 * __shfl_sync(0xffffffffu, x, (base + t + 1) & 31);
    x -= l2 * __shfl_sync(0xffffffffu, x, (base + t + 2) & 31);
    x -= l3 * __shfl_sync(0xffffffffu, x, (base + t + 3) & 31);
    l0 = m0, l1 = m1, l2 = m2, l3 = m3;
  }
  x *= id;
  {
    const int t = 4 * (nblk - 1);
    l0 = (t > tb0 && t < tb1) ? col[t * NB_KLD] : 0.0, l1 = (t + 1 > tb0 && t + 1 < tb1) ? col[(t + 1) * NB_KLD] : 0.0;
    l2 = (t + 2 > tb0 && t + 2 < tb1) ? col[(t + 2) * NB_KLD] : 0.0, l3 = (t + 3 > tb0 && t + 3 < tb1) ? col[(t + 3) * NB_KLD] : 0.0;
  }
#pragma unroll 1
  for (int k = nblk - 1; k >= 0; k--)
  {
    const int t = 4 * k, u = t - 4;
    const double m0 = (u > tb0 && u < tb1) ? col[u * NB_KLD] : 0.0, m1 = (u + 1 > tb0 && u + 1 < tb1) ? col[(u + 1) * NB_KLD] : 0.0,
                 m2 = (u + 2 > tb0 && u + 2 < tb1) ? col[(u + 2) * NB_KLD] : 0.0,
                 m3 = (u + 3 > tb0 && u + 3 < tb1) ? col[(u + 3) * NB_KLD] : 0.0;
    x -= l3 * __shfl_sync(0xffffffffu, x, (base + t + 3) & 31);
    x -= l2 * __shfl_sync(0xffffffffu, x, (base + t + 2) & 31);
    x -= l1 * __shfl_sync(0xffffffffu, x, (base + t + 1) & 31);
    x -= l0 * __shfl_sync(0xffffffffu, x, (base + t) & 31);
    l0 = m0, l1 = m1, l2 = m2, l3 = m3;
  }
  return mine ? x : 0.0;
}
#endif

// Solve K x = b in place, K = K0 (factorised blocks) + dq g g^T by Sherman-Morrison when dq != 0 (u = K0^-1 g and
// gu = g . u precomputed in sh->uq).  Device: x in registers, pivots broadcast by shuffle, the two blocks in lockstep.
template <int NL>
NB_HD void nb_qp_solve_blocks(const Group<NL>& g, NbQpShared* sh, int nxy, int nz, double* b, double dq, double gu,
                               long long* cyc = nullptr)
{
  const int nv = nxy + nz;
  g.sync();
#if defined(__CUDA_ARCH__)
  if (g.lane < 32)
  {
    const int i = g.lane;
    const bool mine = i < nv;
    const long long c0 = cyc ? clock64() : 0;
    double x = nb_warp_solve(sh->K, sh->invd, nxy, nz, i, mine ? b[i] : 0.0);
    if (cyc) *cyc += clock64() - c0;   // measurement hook: the substitution chain alone
    if (dq != 0.0)
    {
      double gy = mine ? sh->gq[i] * x : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gy += __shfl_xor_sync(0xffffffffu, gy, o);
      if (mine) x -= sh->uq[i] * (dq * gy * nb_rcp(1.0 + dq * gu));
    }
    if (mine) b[i] = x;
  }
#else
  const double* K = sh->K;
  for (int blk = 0; blk < 2; blk++)
  {
    const int lo = blk ? nxy : 0, hi = blk ? nv : nxy;
    for (int k = lo; k < hi; k++)
      for (int i = k + 1; i < hi; i++) b[i] -= K[i * NB_KLD + k] * b[k];
    for (int i = lo; i < hi; i++) b[i] *= sh->invd[i];
    for (int k = hi - 1; k >= lo; k--)
      for (int i = lo; i < k; i++) b[i] -= K[k * NB_KLD + i] * b[k];
  }
  if (dq != 0.0)
  {
    double gy = 0.0;
    for (int i = 0; i < nv; i++) gy += sh->gq[i] * b[i];
    for (int i = 0; i < nv; i++) b[i] -= sh->uq[i] * (dq * gy * nb_rcp(1.0 + dq * gu));
  }
#endif
  g.sync();
}

// objective value at the full coefficients (solver_gurobi_poly.cpp:322-380, :882)
NB_HD double nb_objective(const NbConsts& cs, int n, int mode, const double* x /*[3][32]*/, const double pf[3])
{
  const double T = cs.T;
  double f = 0.0;
  for (int ax = 0; ax < 3; ax++)
  {
    for (int i = 0; i < n; i++) f += 36.0 * T * x[ax * 32 + 4 * i] * x[ax * 32 + 4 * i];
    const double* c = x + ax * 32 + 4 * (n - 1);
    const double e = T * T * T * c[0] + T * T * c[1] + T * c[2] + c[3] - pf[ax];
    f += cs.W * e * e;
    if (mode == 1)
    {
      const double v = 3 * T * T * c[0] + 2 * T * c[1] + c[2], a = 6 * T * c[0] + 2 * c[1];
      f += cs.W * (v * v + a * a);
    }
  }
  return f;
}

// Solve one QP attempt.  coeff_init: [3][8][4] of this agent.  On success writes x_out [3][32]
// (coefficients, axis-major) and returns true.  *iters_out = interior-point iterations used.
template <int NL>
NB_HD bool nb_qp_solve(const Group<NL>& g, const NbConsts& cs, const NbQpTable* tb, NbQpShared* sh,
                       const NbQpRows& R, const double* coeff_init, int nlines, double* x_out, int* iters_out,
                       double* obj_out, long long* prof = nullptr)
{
  // measurement hook (nb_set_profiling): SM cycles lane 0 spends per phase, summed over the iterations
#if defined(__CUDA_ARCH__)
  long long pt[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, pc = prof ? clock64() : 0;
#define NB_QP_TICK(k)                     \
  if (prof)                               \
  {                                       \
    const long long now_ = clock64();     \
    pt[k] += now_ - pc;                   \
    pc = now_;                            \
  }
#else
#define NB_QP_TICK(k)
#endif
  const int n = tb->n, dof = tb->dof, nv = 3 * dof, mode = tb->mode;
  const int nxy = 2 * dof, nz = dof;
  const double T = cs.T;
  const double qp[4] = { T * T * T, T * T, T, 1.0 };
  *iters_out = 0;
  // ---- per-agent constants
  g.sync();
  for (int q = g.lane; q < 24; q += NL) nb_feat_bounds(cs, q >> 3, q & 7, sh->blo[q], sh->bhi[q]);
  for (int a = g.lane; a < nv; a += NL)
  {
    sh->ax_of[a] = (unsigned char)(a / (dof > 0 ? dof : 1));
    sh->c_of[a] = (unsigned char)(a - (a / (dof > 0 ? dof : 1)) * dof);
  }
  for (int ca = g.lane; ca < dof; ca += NL)
    for (int cb = 0; cb <= ca; cb++)
    {
      sh->pca[ca * (ca + 1) / 2 + cb] = (unsigned char)ca;
      sh->pcb[ca * (ca + 1) / 2 + cb] = (unsigned char)cb;
    }
  {  // strictly-lower elements of the (x,y) block, then of the z block, row by row
    const int na = nxy * (nxy - 1) / 2, nb = nz * (nz - 1) / 2;
    for (int e = g.lane; e < na + nb; e += NL)
    {
      const int q = e < na ? e : e - na, off = e < na ? 0 : nxy;
      int i = 1;
      while (i * (i + 1) / 2 <= q) i++;
      sh->ei[e] = (unsigned char)(off + i);
      sh->ej[e] = (unsigned char)(off + q - i * (i - 1) / 2);
    }
    if (g.lane == 0) sh->nelem = na + nb;
  }
  for (int q = g.lane; q < NB_NFEAT_AX * 4; q += NL) sh->omv[q] = 0.0;
  for (int q = g.lane; q < 3 * 4 * n; q += NL)
  {
    const int ax = q / (4 * n), r = q - ax * 4 * n;
    sh->xin[ax][r] = coeff_init[ax * 32 + r];
  }
  if (g.lane == 0)
    for (int ax = 0; ax < 3; ax++)
    {
      const double* c0 = coeff_init + ax * 32;
      sh->init3[ax][0] = c0[1];
      sh->init3[ax][1] = c0[2];
      sh->init3[ax][2] = c0[3];
      const double* cl = coeff_init + ax * 32 + 4 * (n - 1);
      sh->pf[ax] = qp[0] * cl[0] + qp[1] * cl[1] + qp[2] * cl[2] + qp[3] * cl[3];  // final_pos_ :226-228
    }
  g.sync();
  const double ddx = sh->init3[0][2] - sh->pf[0], ddy = sh->init3[1][2] - sh->pf[1], ddz = sh->init3[2][2] - sh->pf[2];
  const bool has_qc = sqrt(ddx * ddx + ddy * ddy + ddz * ddz) < 1.0;  // :697-702
  const int m = 48 * n + 4 * nlines, mq = m + (has_qc ? 1 : 0);
  const double inv_mq = 1.0 / mq;
  double bmax = 0.0, hn = 0.0;
  for (int ax = 0; ax < 3; ax++)
    for (int c = 0; c < 3; c++) bmax = fmax(bmax, fabs(sh->init3[ax][c]));
  for (int ax = 0; ax < 3; ax++) hn = fmax(hn, fmax(fabs(cs.lim_min[ax]), fabs(cs.lim_max[ax])));
  hn = fmax(hn, fmax(cs.v_max, cs.a_max));
  {
    double hl = 0.0;
    for (int i = 0; i < n; i++)
      for (int l = R.lstart[i] + g.lane; l < R.lstart[i + 1]; l += NL) hl = fmax(hl, fabs(R.clb[i][3 * l + 2]));
    hn = fmax(hn, g.max(hl));
  }
  if (tb->has_resid)
  {  // n = 1 with terminal v/a equalities: 5 rows on 4 unknowns per axis
    double r = 0.0;
    for (int ax = 0; ax < 3; ax++)
      for (int q = 0; q < 2; q++)
        r = fmax(r, fabs(tb->Rres[q][0] * sh->init3[ax][0] + tb->Rres[q][1] * sh->init3[ax][1] +
                         tb->Rres[q][2] * sh->init3[ax][2]));
    if (r > 1e-9 * (1.0 + bmax)) return false;
  }
  // w0 = Z^T (x_frontend - Pm init3), g0 = Gr init3 + gpf pf
  for (int a = g.lane; a < nv; a += NL)
  {
    const int ax = sh->ax_of[a], c = sh->c_of[a];
    double v = 0.0;
    for (int r = 0; r < 4 * n; r++)
    {
      const double xp = tb->Pm[r][0] * sh->init3[ax][0] + tb->Pm[r][1] * sh->init3[ax][1] + tb->Pm[r][2] * sh->init3[ax][2];
      v += tb->Z[r][c] * (sh->xin[ax][r] - xp);
    }
    sh->w[a] = v;
    sh->g0[a] = tb->Gr[c][0] * sh->init3[ax][0] + tb->Gr[c][1] * sh->init3[ax][1] + tb->Gr[c][2] * sh->init3[ax][2] +
                tb->gpf[c] * sh->pf[ax];
    sh->dw[a] = 0.0;
  }
  // objective at w = 0 (constant term of the reduced objective): lanes rebuild x_p cooperatively
  for (int q = g.lane; q < 3 * 4 * n; q += NL)
  {
    const int ax = q / (4 * n), r = q - ax * 4 * n;
    x_out[ax * 32 + r] = tb->Pm[r][0] * sh->init3[ax][0] + tb->Pm[r][1] * sh->init3[ax][1] + tb->Pm[r][2] * sh->init3[ax][2];
  }
  g.sync();
  const double fconst = nb_objective(cs, n, mode, x_out, sh->pf);
  nb_qp_features<NL>(g, tb, sh, sh->y, sh->w, true);
  g.sync();

#define NB_QC_EVAL(cval)                                                                                   \
  {                                                                                                        \
    cval = -0.10 * 0.10;                                                                                   \
    for (int ax = 0; ax < 3; ax++)                                                                         \
    {                                                                                                      \
      double e = tb->tq0[0] * sh->init3[ax][0] + tb->tq0[1] * sh->init3[ax][1] + tb->tq0[2] * sh->init3[ax][2] - \
                 sh->pf[ax];                                                                               \
      for (int c = 0; c < dof; c++) e += tb->tq[c] * sh->w[ax * dof + c];                                  \
      e3[ax] = e;                                                                                          \
      cval += e * e;                                                                                       \
    }                                                                                                      \
  }

  double e3[3] = { 0, 0, 0 };
  bool converged = false;
  if (dof == 0)
  {  // the equalities leave a single point: feasible iff every row holds there
    double worst = -1e300;
    for (int q = g.lane; q < NB_NF3; q += NL)
    {
      const int ax = q >> 6, fl = q & 63;
      if (fl >= 8 * n) continue;
      const double lo = sh->blo[ax * 8 + (fl & 7)], hi = sh->bhi[ax * 8 + (fl & 7)];
      const double yv = sh->y[q];
      worst = fmax(worst, fmax(yv - hi, lo - yv));
    }
    for (int q = g.lane; q < 4 * n; q += NL)
    {
      const int i = q >> 2, k = q & 3;
      const double* cl = R.clb[i];
      for (int l = R.lstart[i]; l < R.lstart[i + 1]; l++)
        worst = fmax(worst, cl[3 * l] * sh->y[i * 8 + k] + cl[3 * l + 1] * sh->y[NB_NFEAT_AX + i * 8 + k] - cl[3 * l + 2]);
    }
    worst = g.max(worst);
    bool ok = !(worst > 1e-9 * (1.0 + hn));
    if (has_qc)
    {
      double cval;
      NB_QC_EVAL(cval);
      if (cval > 1e-9) ok = false;
    }
    converged = ok;
  }
  else
  {
    // ---- start: s = max(h - G w, 1), lam = 1
    for (int q = g.lane; q < NB_NF3; q += NL)
    {
      const int ax = q >> 6, fl = q & 63, f = q;
      if (fl >= 8 * n) continue;
      const double lo = sh->blo[ax * 8 + (fl & 7)], hi = sh->bhi[ax * 8 + (fl & 7)];
      const double yv = sh->y[f];
      R.s[f] = fmax(hi - yv, 1.0);
      R.s[NB_NF3 + f] = fmax(yv - lo, 1.0);
      R.lam[f] = 1.0;
      R.lam[NB_NF3 + f] = 1.0;
      sh->dy[f] = 0.0;
    }
    for (int q = g.lane; q < 4 * n; q += NL)
    {
      const int i = q >> 2, k = q & 3;
      const double* cl = R.clb[i];
      for (int l = R.lstart[i]; l < R.lstart[i + 1]; l++)
      {
        const int r = NB_ROW_LINE0 + 4 * l + k;
        const double v = cl[3 * l + 2] - cl[3 * l] * sh->y[i * 8 + k] - cl[3 * l + 1] * sh->y[NB_NFEAT_AX + i * 8 + k];
        R.s[r] = fmax(v, 1.0);
        R.lam[r] = 1.0;
      }
    }
    double s_q = 1.0, lam_q = 1.0, dsa_q = 0.0, dla_q = 0.0;
    if (has_qc)
    {
      double cval;
      NB_QC_EVAL(cval);
      s_q = fmax(-cval, 1.0);
    }
    g.sync();

    NB_QP_TICK(0);
    int it = 0;
    double al_pending = 0.0;  // step of the previous iteration, applied to (s, lam, y) by the next RESID sweep
    double mu_div = 1e300;    // divergence threshold, set from the first real iterate
    for (it = 0; it <= cs.max_iter; it++)
    {
      const bool init_pass = (it == 0);
      // ---- residuals and weights; beside them gobj = Hr w + g0 (the sweep's last barrier publishes both)
      double cval = 0.0, rp_q = 0.0;
      if (has_qc)
      {
        NB_QC_EVAL(cval);
        rp_q = cval + s_q;
      }
      for (int a = g.lane; a < nv; a += NL)
      {
        const int ax = sh->ax_of[a], c = sh->c_of[a];
        double go = sh->g0[a];
#pragma unroll
        for (int c2 = 0; c2 < NB_DOF_MAX; c2++)
          if (c2 < dof) go += tb->Hr[c][c2] * sh->w[ax * dof + c2];   // unrolled: the loads go out together
        sh->gobj[a] = go;
        sh->gq[a] = has_qc ? 2.0 * e3[ax] * tb->tq[c] : 0.0;
      }
      NbPassAcc acc = { 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_RESID>(g, tb, sh, R, nlines, 0.0, al_pending, acc);
      al_pending = 0.0;
      NB_QP_TICK(1);
      double sum_sl = acc.sum_sl, rp_max = acc.rp_max, unused_ = 0.0;
      g.get(0, sum_sl, rp_max, unused_);
      const double mu = (sum_sl + (has_qc ? s_q * lam_q : 0.0)) * inv_mq;
      double rpn = rp_max;
      if (has_qc) rpn = fmax(rpn, fabs(rp_q));
      // r_d = gobj + C^T(lambda load) + lam_q grad c
      nb_qp_ct_all<NL>(g, tb, sh, sh->du, sh->rd, sh->gobj, 1.0, 1.0, sh->gq, has_qc ? lam_q : 0.0);
      g.sync();
      NB_QP_TICK(11);
      double rdn = 0.0, gn = 0.0, fobj = fconst;
#if defined(__CUDA_ARCH__)
      {  // nv <= 24 values in shared memory: every WARP folds them itself, one value per lane (no barrier)
        const int a = g.lane & 31;
        const bool on = a < nv;
        const double go = on ? sh->gobj[a] : 0.0;
        double fo = on ? 0.5 * (sh->g0[a] + go) * sh->w[a] : 0.0;
        rdn = on ? fabs(sh->rd[a]) : 0.0, gn = fabs(go);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          rdn = fmax(rdn, __shfl_xor_sync(0xffffffffu, rdn, o));
          gn = fmax(gn, __shfl_xor_sync(0xffffffffu, gn, o));
          fo += __shfl_xor_sync(0xffffffffu, fo, o);
        }
        fobj += fo;
      }
#else
      for (int a = 0; a < nv; a++)
      {
        rdn = fmax(rdn, fabs(sh->rd[a]));
        gn = fmax(gn, fabs(sh->gobj[a]));
        fobj += 0.5 * (sh->g0[a] + sh->gobj[a]) * sh->w[a];
      }
#endif
      if (!init_pass)
      {
        if (rpn <= cs.tol * (1.0 + hn) && rdn <= cs.tol * (1.0 + gn) && mu * mq <= cs.tol * (1.0 + fabs(fobj)))
        {
          converged = true;
          break;
        }
        if (it == cs.max_iter) break;
        if (!(mu == mu) || !(rpn == rpn) || !(rdn == rdn)) break;
        // diverged: on an infeasible model the multipliers run away (mu grows by ~1e5 per iteration while the steps
        // collapse to 1e-50) long before anything overflows to NaN; a converging solve never leaves mu twelve orders of
        // magnitude ABOVE where it started.  Same outcome as the NaN test and the iteration cap ("no solution"), sooner.
        if (it == 1) mu_div = 1e12 * (1.0 + mu);
        if (mu > mu_div) break;
      }
      NB_QP_TICK(2);
      // ---- K0 = Hobj + C^T W C (+ Hessian of the quadratic row), factorised block by block
      nb_qp_assemble<NL>(g, tb, sh, has_qc, lam_q);
      NB_QP_TICK(3);
      nb_qp_factor_blocks<NL>(g, sh, nxy, nz);
      NB_QP_TICK(4);
      const double is_q = has_qc ? nb_rcp(s_q) : 0.0, il_q = has_qc ? nb_rcp(lam_q) : 0.0;
      const double d_q = lam_q * is_q;
      double gu = 0.0;
      if (has_qc)
      {  // Sherman-Morrison for the rank-one term d_q gq gq^T: u = K0^-1 gq
        for (int a = g.lane; a < nv; a += NL) sh->uq[a] = sh->gq[a];
        nb_qp_solve_blocks<NL>(g, sh, nxy, nz, sh->uq, 0.0, 0.0);
        double t = 0.0;
        for (int a = g.lane; a < nv; a += NL) t += sh->gq[a] * sh->uq[a];
        gu = g.sum(t);
      }
      // ---- predictor
      const double tau_q = (lam_q * rp_q - s_q * lam_q) * is_q;
      nb_qp_ct_all<NL>(g, tb, sh, sh->La, sh->dw, sh->rd, -1.0, -1.0, sh->gq, -tau_q);
      NB_QP_TICK(9);
#if defined(__CUDA_ARCH__)
      nb_qp_solve_blocks<NL>(g, sh, nxy, nz, sh->dw, d_q, gu, prof ? &pt[12] : nullptr);
#else
      nb_qp_solve_blocks<NL>(g, sh, nxy, nz, sh->dw, d_q, gu);
#endif
      NB_QP_TICK(10);
      nb_qp_features<NL>(g, tb, sh, sh->dy, sh->dw, false);
      g.sync();
      NB_QP_TICK(5);
      acc = NbPassAcc{ 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_DIR_PRED>(g, tb, sh, R, nlines, 0.0, 0.0, acc);
      double rmax = acc.rmax, scross = acc.sum_cross, sdd = acc.sum_dd;
      g.get(1, scross, rmax, sdd);
      NB_QP_TICK(6);
      if (has_qc)
      {
        double gd = 0.0;
        for (int a = 0; a < nv; a++) gd += sh->gq[a] * sh->dw[a];
        dsa_q = -rp_q - gd;
        dla_q = (-s_q * lam_q - lam_q * dsa_q) * is_q;
        rmax = fmax(rmax, fmax(-dsa_q * is_q, -dla_q * il_q));
        scross += s_q * dla_q + lam_q * dsa_q;
        sdd += dsa_q * dla_q;
      }
      if (init_pass)
      {  // Nocedal-Wright starting-point correction
        nb_qp_pass<NL, NB_PASS_START>(g, tb, sh, R, nlines, 0.0, 0.0, acc);
        if (has_qc)
        {
          s_q = fmax(fabs(s_q + dsa_q), 1.0);
          lam_q = fmax(fabs(lam_q + dla_q), 1.0);
        }
        continue;
      }
      const double a_aff = rmax > 1.0 ? nb_rcp(rmax) : 1.0;
      const double mu_aff = (mu * mq + a_aff * scross + a_aff * a_aff * sdd) * inv_mq;
      const double rat = mu_aff * nb_rcp(mu);
      double sigmu = rat * rat * rat * mu;
      // do not drive the complementarity below a tenth of what the stopping test needs: at mu ~ 1e-12 the
      // normal matrix is too ill-conditioned for the dual residual to reach its tolerance
      sigmu = fmax(sigmu, 0.1 * cs.tol * (1.0 + fabs(fobj)) * inv_mq);
      // ---- corrector: load = predictor load - ds_aff dl_aff / s + sigma mu / s, per feature
      for (int q = g.lane; q < NB_NF3; q += NL)
        if ((q & 63) < 8 * n) sh->La[q] += sigmu * sh->A2[q] - sh->A1[q];
      g.sync();
      const double rc_q = s_q * lam_q + dsa_q * dla_q - sigmu;
      const double tau_q2 = (lam_q * rp_q - rc_q) * is_q;
      nb_qp_ct_all<NL>(g, tb, sh, sh->La, sh->dw, sh->rd, -1.0, -1.0, sh->gq, -tau_q2);
      nb_qp_solve_blocks<NL>(g, sh, nxy, nz, sh->dw, d_q, gu);
      nb_qp_features<NL>(g, tb, sh, sh->dy, sh->dw, false);
      g.sync();
      NB_QP_TICK(7);
      acc = NbPassAcc{ 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_DIR_CORR>(g, tb, sh, R, nlines, sigmu, 0.0, acc);
      rmax = acc.rmax;
      g.get(0, unused_, rmax, unused_);
      NB_QP_TICK(8);
      double ds_q = 0.0, dl_q = 0.0;
      if (has_qc)
      {
        double gd = 0.0;
        for (int a = 0; a < nv; a++) gd += sh->gq[a] * sh->dw[a];
        ds_q = -rp_q - gd;
        dl_q = (-rc_q - lam_q * ds_q) * is_q;
        rmax = fmax(rmax, fmax(-ds_q * is_q, -dl_q * il_q));
      }
      const double eta = 1.0 - nb_rcp((it + 3.0) * (it + 3.0));
      double al = rmax > 0.0 ? eta * nb_rcp(rmax) : 1.0;
      if (al > 1.0) al = 1.0;
      al_pending = al;
#if defined(__CUDA_ARCH__)
      if (prof && g.lane == 0 && it >= 16)
      {  // diagnostic (profiling mode): the late iterations of a solve that is not converging -- slots 13..15 hold the
         // smallest step, the last primal residual and the last mu as raw bits
        double* pd = reinterpret_cast<double*>(prof + 13);
        pd[0] = (it == 16 || al < pd[0]) ? al : pd[0];
        pd[1] = rpn;
        pd[2] = mu;
      }
#endif
      if (has_qc)
      {
        s_q += al * ds_q;
        lam_q += al * dl_q;
      }
      for (int a = g.lane; a < nv; a += NL) sh->w[a] += al * sh->dw[a];
      g.sync();
    }
    *iters_out = it;
  }
#if defined(__CUDA_ARCH__)
  if (prof && g.lane == 0)
    for (int k = 0; k < 13; k++) prof[k] += pt[k];
#endif
#undef NB_QP_TICK
#undef NB_QC_EVAL
  if (converged)
  {
    for (int q = g.lane; q < 3 * 4 * n; q += NL)
    {
      const int ax = q / (4 * n), r = q - ax * 4 * n;
      double v = tb->Pm[r][0] * sh->init3[ax][0] + tb->Pm[r][1] * sh->init3[ax][1] + tb->Pm[r][2] * sh->init3[ax][2];
      for (int c = 0; c < dof; c++) v += tb->Z[r][c] * sh->w[ax * dof + c];
      x_out[ax * 32 + r] = v;
    }
    g.sync();
    *obj_out = nb_objective(cs, n, mode, x_out, sh->pf);
  }
  return converged;
}
