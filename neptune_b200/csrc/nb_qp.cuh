// nb_qp.cuh -- K4: batched trajectory QP, one agent per warp-group.
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
// Lane-strided SPMD phases over shared memory (vectors, K, per-row s / lambda; global scratch only when an
// agent keeps more lines than fit); compiles with NL = 128 (one CTA per agent) on the device and NL = 1 in
// the host emulation.
#pragma once
#include "nb_common.cuh"

struct NbQpShared
{
  double y[3 * NB_NFEAT_AX];   // feature values
  double dy[3 * NB_NFEAT_AX];  // feature values of a direction
  double om[3 * NB_NFEAT_AX];  // diagonal weights  sum lambda/s
  double La[3 * NB_NFEAT_AX];  // per-feature load of a row vector (G^T v reduced to features)
  double Sxy[4 * NB_NPOL];     // x-y coupling weight of control point (i,k)
  double K[NB_NV_MAX * NB_NV_MAX];
  double w[NB_NV_MAX], dw[NB_NV_MAX], rd[NB_NV_MAX], rhs[NB_NV_MAX], g0[NB_NV_MAX], gq[NB_NV_MAX];
  double gobj[NB_NV_MAX], invd[NB_NV_MAX];
  double init3[3][3], pf[3], e3[3];
  double xin[3][4 * NB_NPOL];
  double blo[24], bhi[24];                     // bounds of feature kind (axis, j)
  unsigned char ax_of[NB_NV_MAX], c_of[NB_NV_MAX];  // variable a -> (axis, column)
  unsigned char pa[NB_NV_MAX * (NB_NV_MAX + 1) / 2], pb[NB_NV_MAX * (NB_NV_MAX + 1) / 2];  // lower-triangle pairs
  int npairs;
};

NB_HD double nb_rcp(double x)
{
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);  // correctly rounded reciprocal == 1.0 / x, without the division slow path
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
  const double* cl;  // [L][3] compact kept lines: n0, n1, c = 1 - d
  const int* lstart; // [n+1] first line of each interval
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

// y = c0 . init3 + C w  (with_const) or dy = C dw
template <int NL>
NB_HD void nb_qp_features(const Group<NL>& g, const NbQpTable* tb, const NbQpShared* sh_c, double* out,
                          const double* vec, bool with_const)
{
  const int n = tb->n, dof = tb->dof;
  for (int q = g.lane; q < 3 * NB_NFEAT_AX; q += NL)
  {
    const int ax = q >> 6, fl = q & 63;
    if (fl >= 8 * n) continue;
    double v = 0.0;
    if (with_const)
      v = tb->c0[fl][0] * sh_c->init3[ax][0] + tb->c0[fl][1] * sh_c->init3[ax][1] + tb->c0[fl][2] * sh_c->init3[ax][2];
    for (int c = 0; c < dof; c++) v += tb->C[fl][c] * vec[ax * dof + c];
    out[ax * NB_NFEAT_AX + fl] = v;
  }
}

enum
{
  NB_PASS_RESID = 0,  // residuals, weights, predictor load; sums mu, |rp|max
  NB_PASS_DIR_PRED,   // predictor direction: ds, dl -> scratch; step length, mu_aff sums
  NB_PASS_LOAD_CORR,  // corrector load with rc = s lam + dsa dla - sigma mu
  NB_PASS_DIR_CORR,   // final direction: ds, dl -> scratch; step length
  NB_PASS_UPDATE,     // s += a ds, lam += a dl
  NB_PASS_START       // Nocedal-Wright start: s = max(1,|s+ds|), lam likewise
};

struct NbPassAcc
{
  double sum_sl, rp_max, rmax, sum_cross, sum_dd;  // rmax = max over rows of (-ds/s, -dl/lam): alpha_max = 1/rmax
};

// one inequality row: v = row value (<= 0 wanted), gd = row . direction.  Returns the load tau that
// this row puts on its feature(s) (for RESID / LOAD_CORR), and the diagonal weight in `wgt`.
template <int MODE>
NB_HD double nb_row(double* s_, double* lam_, double* dsa_, double* dla_, double* inv_, double v, double gd,
                    double sigmu, double alpha, NbPassAcc& acc, double& wgt)
{
  double s = *s_, lam = *lam_;
  wgt = 0.0;
  if (MODE == NB_PASS_RESID)
  {
    if (alpha != 0.0)
    {  // pending step of the previous iteration (s, lam) += alpha (ds, dl); v already is at the new point
      s += alpha * (*dsa_);
      lam += alpha * (*dla_);
      *s_ = s;
      *lam_ = lam;
    }
    const double rp = v + s;
    const double inv = nb_rcp(s);
    *inv_ = inv;
    wgt = lam * inv;
    acc.sum_sl += s * lam;
    acc.rp_max = fmax(acc.rp_max, fabs(rp));
    return (lam * rp - s * lam) * inv;  // predictor: rc = s lam
  }
  if (MODE == NB_PASS_DIR_PRED)
  {
    const double inv = *inv_;
    const double ds = -(v + s) - gd;
    const double dl = (-s * lam - lam * ds) * inv;
    acc.rmax = fmax(acc.rmax, -ds * inv);       // -ds/s
    acc.rmax = fmax(acc.rmax, 1.0 + ds * inv);  // -dl/lam = (s + ds)/s for the predictor
    acc.sum_cross += s * dl + lam * ds;
    acc.sum_dd += ds * dl;
    *dsa_ = ds;
    *dla_ = dl;
    return 0.0;
  }
  if (MODE == NB_PASS_DIR_CORR)
  {
    const double inv = *inv_;
    const double rc = s * lam + (*dsa_) * (*dla_) - sigmu;
    const double ds = -(v + s) - gd;
    const double dl = (-rc - lam * ds) * inv;
    acc.rmax = fmax(acc.rmax, -ds * inv);
    acc.rmax = fmax(acc.rmax, -dl * nb_rcp(lam));
    *dsa_ = ds;
    *dla_ = dl;
    return 0.0;
  }
  if (MODE == NB_PASS_LOAD_CORR)
  {
    const double rp = v + s;
    const double rc = s * lam + (*dsa_) * (*dla_) - sigmu;
    return (lam * rp - rc) * (*inv_);
  }
  if (MODE == NB_PASS_UPDATE)
  {
    *s_ = s + alpha * (*dsa_);
    *lam_ = lam + alpha * (*dla_);
    return 0.0;
  }
  // NB_PASS_START
  {
    const double a = fabs(s + *dsa_), b = fabs(lam + *dla_);
    *s_ = a > 1.0 ? a : 1.0;
    *lam_ = b > 1.0 ? b : 1.0;
  }
  return 0.0;
}

// One sweep over every inequality row.  Bound rows are visited feature by feature (both sides of a
// feature by the same lane), line rows control point by control point: item (interval i, control
// point k, sub-lane) where the SUB adjacent lanes of an item split its lines and combine by shuffle.
// All per-feature accumulations are conflict-free and in a fixed order.
template <int NL, int MODE>
NB_HD void nb_qp_pass(const Group<NL>& g, const NbConsts& cs, const NbQpTable* tb, NbQpShared* sh,
                      const NbQpRows& R, double sigmu, double alpha, NbPassAcc& acc)
{
  constexpr int SUB = Group<NL>::SUB;
  const int n = tb->n;
  const bool loads = (MODE == NB_PASS_RESID || MODE == NB_PASS_LOAD_CORR);
  for (int q = g.lane; q < 3 * NB_NFEAT_AX; q += NL)
  {
    const int ax = q >> 6, fl = q & 63, f = q;
    if (fl >= 8 * n) continue;
    double w_u, w_l;
    const double lo = sh->blo[ax * 8 + (fl & 7)], hi = sh->bhi[ax * 8 + (fl & 7)];
    double yv = sh->y[f];
    const double dv = sh->dy[f];
    if (MODE == NB_PASS_RESID && alpha != 0.0)
    {  // features are linear in w: y(w + alpha dw) = y + alpha dy
      yv += alpha * dv;
      sh->y[f] = yv;
    }
    const int rs = 2 * f;
    const double t_u = nb_row<MODE>(R.s + rs, R.lam + rs, R.dsa + rs, R.dla + rs, R.inv + rs, yv - hi, dv, sigmu, alpha,
                                    acc, w_u);
    const double t_l = nb_row<MODE>(R.s + rs + 1, R.lam + rs + 1, R.dsa + rs + 1, R.dla + rs + 1, R.inv + rs + 1,
                                    lo - yv, -dv, sigmu, alpha, acc, w_l);
    if (loads)
    {
      sh->La[f] = t_u - t_l;
      if (MODE == NB_PASS_RESID)
      {
        sh->om[f] = w_u + w_l;
        sh->dy[f] = R.lam[rs] - R.lam[rs + 1];  // dual load for r_d, parked in dy during RESID
      }
    }
  }
  g.sync();
  const int items = 4 * n * SUB;  // <= NL for SUB > 1, so every lane reaches the shuffles below
  for (int base = 0; base < items; base += NL)
  {
    const int q = base + g.lane;
    const bool on = q < items;
    const int gq = on ? q / SUB : 0, sub = q % SUB;
    const int i = gq >> 2, k = gq & 3;
    const int fx = i * 8 + k, fy = NB_NFEAT_AX + i * 8 + k;
    const double yx = sh->y[fx], yy = sh->y[fy];
    const double dx = (MODE == NB_PASS_RESID) ? 0.0 : sh->dy[fx], dyv = (MODE == NB_PASS_RESID) ? 0.0 : sh->dy[fy];
    double sxx = 0, sxy = 0, syy = 0, lx = 0, ly = 0, dlx = 0, dly = 0;
    if (on)
      for (int l = R.lstart[i] + sub; l < R.lstart[i + 1]; l += SUB)
      {
        const double n0 = R.cl[3 * l], n1 = R.cl[3 * l + 1], c = R.cl[3 * l + 2];
        const int rs = 6 * NB_NFEAT_AX + 4 * l + k;
        double wgt;
        const double t = nb_row<MODE>(R.s + rs, R.lam + rs, R.dsa + rs, R.dla + rs, R.inv + rs, n0 * yx + n1 * yy - c,
                                      n0 * dx + n1 * dyv, sigmu, alpha, acc, wgt);
        if (loads)
        {
          lx += t * n0;
          ly += t * n1;
          if (MODE == NB_PASS_RESID)
          {
            sxx += wgt * n0 * n0;
            sxy += wgt * n0 * n1;
            syy += wgt * n1 * n1;
            const double lam = R.lam[rs];
            dlx += lam * n0;
            dly += lam * n1;
          }
        }
      }
    if (loads)
    {
      lx = g.sub_sum(lx), ly = g.sub_sum(ly);
      if (MODE == NB_PASS_RESID)
      {
        sxx = g.sub_sum(sxx), sxy = g.sub_sum(sxy), syy = g.sub_sum(syy), dlx = g.sub_sum(dlx), dly = g.sub_sum(dly);
      }
      if (on && sub == 0)
      {
        sh->La[fx] += lx;
        sh->La[fy] += ly;
        if (MODE == NB_PASS_RESID)
        {
          sh->om[fx] += sxx;
          sh->om[fy] += syy;
          sh->Sxy[gq] = sxy;
          sh->dy[fx] += dlx;
          sh->dy[fy] += dly;
        }
      }
    }
  }
  g.sync();
}

// out[a] = sum_f C[f][c] * vecF[ax*64+f]   (a = ax*dof + c): C^T applied to a per-feature vector, then
// dst[a] = bsign * base[a] + sign * that + es * extra[a]; item (a, sub-lane) with shuffle combine
template <int NL>
NB_HD void nb_qp_ct_all(const Group<NL>& g, const NbQpTable* tb, const NbQpShared* sh_t, const double* vecF, double* dst,
                        const double* base, double bsign, double sign, const double* extra, double es)
{
  constexpr int SUB = Group<NL>::SUB;
  const int n = tb->n, dof = tb->dof, nv = 3 * dof;
  const int items = nv * SUB;
  for (int b0 = 0; b0 < items; b0 += NL)
  {
    const int q = b0 + g.lane;
    const bool on = q < items;
    const int a = on ? q / SUB : 0, sub = q % SUB;
    const int ax = sh_t->ax_of[a], c = sh_t->c_of[a];
    double v0 = 0.0, v1 = 0.0;
    if (on)
    {
      const double* vf = vecF + ax * NB_NFEAT_AX;
      int fl = sub;
      for (; fl + SUB < 8 * n; fl += 2 * SUB)
      {
        v0 += tb->C[fl][c] * vf[fl];
        v1 += tb->C[fl + SUB][c] * vf[fl + SUB];
      }
      if (fl < 8 * n) v0 += tb->C[fl][c] * vf[fl];
    }
    const double v = g.sub_sum(v0 + v1);
    if (on && sub == 0) dst[a] = bsign * base[a] + sign * v + (extra ? extra[a] * es : 0.0);
  }
}

// Factorisation K = L D L^T of the nv x nv matrix in sh->K (lower triangle, row-major, ld = nv):
// afterwards K[i][k] (i > k) = L[i][k] and invd[k] = 1 / D[k].  sh->rhs is used as a column scratch.
template <int NL>
NB_HD void nb_qp_factor(const Group<NL>& g, NbQpShared* sh, int nv)
{
  double* K = sh->K;
  double* col = sh->rhs;
  for (int k = 0; k < nv; k++)
  {
    double d = K[k * nv + k];
    d = d > 1e-300 ? d : 1e-300;  // rank-deficiency guard (K is positive definite by construction)
    const double ip = nb_rcp(d);
    if (g.lane == 0) sh->invd[k] = ip;
    for (int i = k + 1 + g.lane; i < nv; i += NL)
    {
      const double t = K[i * nv + k];
      col[i] = t;
      K[i * nv + k] = t * ip;
    }
    g.sync();
    for (int q = g.lane; q < sh->npairs; q += NL)
    {
      const int i = sh->pa[q], j = sh->pb[q];
      if (j > k) K[i * nv + j] -= K[i * nv + k] * col[j];
    }
    g.sync();
  }
}

// Solve L D L^T x = b in place.  Device: one warp, x in registers, broadcasts by shuffle (no block
// barriers); host emulation: plain loops in the same summation order.
template <int NL>
NB_HD void nb_qp_solve_ldl(const Group<NL>& g, NbQpShared* sh, int nv, double* b)
{
  const double* K = sh->K;
  g.sync();
#if defined(__CUDA_ARCH__)
  if (g.lane < 32)
  {
    const int i = g.lane;
    double x = i < nv ? b[i] : 0.0;
    for (int k = 0; k < nv; k++)
    {  // forward: unit lower
      const double xk = __shfl_sync(0xffffffffu, x, k);
      if (i > k && i < nv) x -= K[i * nv + k] * xk;
    }
    if (i < nv) x *= sh->invd[i];
    for (int k = nv - 1; k >= 0; k--)
    {  // backward: L^T
      const double xk = __shfl_sync(0xffffffffu, x, k);
      if (i < k) x -= K[k * nv + i] * xk;
    }
    if (i < nv) b[i] = x;
  }
#else
  for (int k = 0; k < nv; k++)
    for (int i = k + 1; i < nv; i++) b[i] -= K[i * nv + k] * b[k];
  for (int i = 0; i < nv; i++) b[i] *= sh->invd[i];
  for (int k = nv - 1; k >= 0; k--)
    for (int i = 0; i < k; i++) b[i] -= K[k * nv + i] * b[k];
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
                       double* obj_out)
{
  const int n = tb->n, dof = tb->dof, nv = 3 * dof, mode = tb->mode;
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
  if (g.lane == 0)
  {
    int q = 0;
    for (int a = 0; a < nv; a++)
      for (int b = 0; b <= a; b++)
      {
        sh->pa[q] = (unsigned char)a;
        sh->pb[q] = (unsigned char)b;
        q++;
      }
    sh->npairs = q;
  }
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
  double bmax = 0.0, hn = 0.0;
  for (int ax = 0; ax < 3; ax++)
    for (int c = 0; c < 3; c++) bmax = fmax(bmax, fabs(sh->init3[ax][c]));
  for (int ax = 0; ax < 3; ax++) hn = fmax(hn, fmax(fabs(cs.lim_min[ax]), fabs(cs.lim_max[ax])));
  hn = fmax(hn, fmax(cs.v_max, cs.a_max));
  {
    double hl = 0.0;
    for (int l = g.lane; l < nlines; l += NL) hl = fmax(hl, fabs(R.cl[3 * l + 2]));
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
    for (int q = g.lane; q < 3 * NB_NFEAT_AX; q += NL)
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
      for (int l = R.lstart[i]; l < R.lstart[i + 1]; l++)
        worst = fmax(worst, R.cl[3 * l] * sh->y[i * 8 + k] + R.cl[3 * l + 1] * sh->y[NB_NFEAT_AX + i * 8 + k] - R.cl[3 * l + 2]);
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
    for (int q = g.lane; q < 3 * NB_NFEAT_AX; q += NL)
    {
      const int ax = q >> 6, fl = q & 63, f = q;
      if (fl >= 8 * n) continue;
      const double lo = sh->blo[ax * 8 + (fl & 7)], hi = sh->bhi[ax * 8 + (fl & 7)];
      const double yv = sh->y[f];
      R.s[2 * f] = fmax(hi - yv, 1.0);
      R.s[2 * f + 1] = fmax(yv - lo, 1.0);
      R.lam[2 * f] = 1.0;
      R.lam[2 * f + 1] = 1.0;
    }
    for (int q = g.lane; q < 4 * nlines; q += NL)
    {
      const int l = q >> 2, k = q & 3;
      int i = 0;
      while (l >= R.lstart[i + 1]) i++;
      const int rs = 6 * NB_NFEAT_AX + q;
      const double v = R.cl[3 * l + 2] - R.cl[3 * l] * sh->y[i * 8 + k] - R.cl[3 * l + 1] * sh->y[NB_NFEAT_AX + i * 8 + k];
      R.s[rs] = fmax(v, 1.0);
      R.lam[rs] = 1.0;
    }
    double s_q = 1.0, lam_q = 1.0, dsa_q = 0.0, dla_q = 0.0;
    if (has_qc)
    {
      double cval;
      NB_QC_EVAL(cval);
      s_q = fmax(-cval, 1.0);
    }
    g.sync();

    int it = 0;
    double al_pending = 0.0;  // step of the previous iteration, applied to (s, lam, y) by the next RESID sweep
    for (it = 0; it <= cs.max_iter; it++)
    {
      const bool init_pass = (it == 0);
      // ---- residuals and weights
      NbPassAcc acc = { 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_RESID>(g, cs, tb, sh, R, 0.0, al_pending, acc);
      al_pending = 0.0;
      double cval = 0.0, rp_q = 0.0;
      if (has_qc)
      {
        NB_QC_EVAL(cval);
        rp_q = cval + s_q;
      }
      double d0 = 0.0, d1 = 0.0;
      g.reduce5(acc.sum_sl, acc.rp_max, acc.rmax, d0, d1);
      const double mu = (acc.sum_sl + (has_qc ? s_q * lam_q : 0.0)) / mq;
      double rpn = acc.rp_max;
      if (has_qc) rpn = fmax(rpn, fabs(rp_q));
      // gobj = Hr w + g0 ; r_d = gobj + C^T(lambda load) + lam_q grad c
      for (int a = g.lane; a < nv; a += NL)
      {
        const int ax = sh->ax_of[a], c = sh->c_of[a];
        double go = sh->g0[a];
        for (int c2 = 0; c2 < dof; c2++) go += tb->Hr[c][c2] * sh->w[ax * dof + c2];
        sh->gobj[a] = go;
        sh->gq[a] = has_qc ? 2.0 * e3[ax] * tb->tq[c] : 0.0;
      }
      g.sync();
      nb_qp_ct_all<NL>(g, tb, sh, sh->dy, sh->rd, sh->gobj, 1.0, 1.0, sh->gq, has_qc ? lam_q : 0.0);
      g.sync();
      double rdn = 0.0, gn = 0.0, fobj = fconst;
      for (int a = 0; a < nv; a++)
      {
        rdn = fmax(rdn, fabs(sh->rd[a]));
        gn = fmax(gn, fabs(sh->gobj[a]));
        fobj += 0.5 * (sh->g0[a] + sh->gobj[a]) * sh->w[a];
      }
      if (!init_pass)
      {
        if (rpn <= cs.tol * (1.0 + hn) && rdn <= cs.tol * (1.0 + gn) && mu * mq <= cs.tol * (1.0 + fabs(fobj)))
        {
          converged = true;
          break;
        }
        if (it == cs.max_iter) break;
        if (!(mu == mu) || !(rpn == rpn) || !(rdn == rdn)) break;
      }
      // ---- K = Hobj + C^T W C (+ quadratic-constraint terms), lower triangle
      const double d_q = has_qc ? lam_q / s_q : 0.0;
      for (int q = g.lane; q < sh->npairs; q += NL)
      {
        const int a = sh->pa[q], b = sh->pb[q];
        const int axa = sh->ax_of[a], ca = sh->c_of[a], axb = sh->ax_of[b], cb = sh->c_of[b];
        double v = 0.0;
        if (axa == axb)
        {
          const double* om = sh->om + axa * NB_NFEAT_AX;
          double v0 = tb->Hr[ca][cb], v1 = 0.0;
          for (int fl = 0; fl < 8 * n; fl += 2)
          {
            v0 += om[fl] * tb->C[fl][ca] * tb->C[fl][cb];
            v1 += om[fl + 1] * tb->C[fl + 1][ca] * tb->C[fl + 1][cb];
          }
          v = v0 + v1;
          if (has_qc) v += lam_q * 2.0 * tb->tq[ca] * tb->tq[cb];
        }
        else if (axa == 1 && axb == 0)
        {
          double v0 = 0.0, v1 = 0.0;
          for (int i = 0; i < n; i++)
          {
            v0 += sh->Sxy[i * 4] * tb->C[i * 8][ca] * tb->C[i * 8][cb] +
                  sh->Sxy[i * 4 + 2] * tb->C[i * 8 + 2][ca] * tb->C[i * 8 + 2][cb];
            v1 += sh->Sxy[i * 4 + 1] * tb->C[i * 8 + 1][ca] * tb->C[i * 8 + 1][cb] +
                  sh->Sxy[i * 4 + 3] * tb->C[i * 8 + 3][ca] * tb->C[i * 8 + 3][cb];
          }
          v = v0 + v1;
        }
        if (has_qc) v += d_q * sh->gq[a] * sh->gq[b];
        sh->K[a * nv + b] = v;
      }
      g.sync();
      nb_qp_factor<NL>(g, sh, nv);
      // ---- predictor
      const double tau_q = has_qc ? (lam_q * rp_q - s_q * lam_q) / s_q : 0.0;
      nb_qp_ct_all<NL>(g, tb, sh, sh->La, sh->dw, sh->rd, -1.0, -1.0, sh->gq, -tau_q);
      nb_qp_solve_ldl<NL>(g, sh, nv, sh->dw);
      nb_qp_features<NL>(g, tb, sh, sh->dy, sh->dw, false);
      g.sync();
      acc = NbPassAcc{ 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_DIR_PRED>(g, cs, tb, sh, R, 0.0, 0.0, acc);
      g.reduce5(d0, acc.rmax, d1, acc.sum_cross, acc.sum_dd);
      double rmax = acc.rmax, scross = acc.sum_cross, sdd = acc.sum_dd;
      if (has_qc)
      {
        double gd = 0.0;
        for (int a = 0; a < nv; a++) gd += sh->gq[a] * sh->dw[a];
        dsa_q = -rp_q - gd;
        dla_q = (-s_q * lam_q - lam_q * dsa_q) / s_q;
        rmax = fmax(rmax, fmax(-dsa_q / s_q, -dla_q / lam_q));
        scross += s_q * dla_q + lam_q * dsa_q;
        sdd += dsa_q * dla_q;
      }
      if (init_pass)
      {  // Nocedal-Wright starting-point correction
        nb_qp_pass<NL, NB_PASS_START>(g, cs, tb, sh, R, 0.0, 0.0, acc);
        if (has_qc)
        {
          s_q = fmax(fabs(s_q + dsa_q), 1.0);
          lam_q = fmax(fabs(lam_q + dla_q), 1.0);
        }
        continue;
      }
      const double a_aff = rmax > 1.0 ? 1.0 / rmax : 1.0;
      const double mu_aff = (mu * mq + a_aff * scross + a_aff * a_aff * sdd) / mq;
      const double rat = mu_aff / mu;
      double sigmu = rat * rat * rat * mu;
      // do not drive the complementarity below a tenth of what the stopping test needs: at mu ~ 1e-12 the
      // normal matrix is too ill-conditioned for the dual residual to reach its tolerance
      sigmu = fmax(sigmu, 0.1 * cs.tol * (1.0 + fabs(fobj)) / mq);
      // ---- corrector
      nb_qp_pass<NL, NB_PASS_LOAD_CORR>(g, cs, tb, sh, R, sigmu, 0.0, acc);
      const double rc_q = s_q * lam_q + dsa_q * dla_q - sigmu;
      const double tau_q2 = has_qc ? (lam_q * rp_q - rc_q) / s_q : 0.0;
      nb_qp_ct_all<NL>(g, tb, sh, sh->La, sh->dw, sh->rd, -1.0, -1.0, sh->gq, -tau_q2);
      nb_qp_solve_ldl<NL>(g, sh, nv, sh->dw);
      nb_qp_features<NL>(g, tb, sh, sh->dy, sh->dw, false);
      g.sync();
      acc = NbPassAcc{ 0.0, 0.0, 0.0, 0.0, 0.0 };
      nb_qp_pass<NL, NB_PASS_DIR_CORR>(g, cs, tb, sh, R, sigmu, 0.0, acc);
      rmax = g.max(acc.rmax);
      double ds_q = 0.0, dl_q = 0.0;
      if (has_qc)
      {
        double gd = 0.0;
        for (int a = 0; a < nv; a++) gd += sh->gq[a] * sh->dw[a];
        ds_q = -rp_q - gd;
        dl_q = (-rc_q - lam_q * ds_q) / s_q;
        rmax = fmax(rmax, fmax(-ds_q / s_q, -dl_q / lam_q));
      }
      const double eta = 1.0 - 1.0 / ((it + 3.0) * (it + 3.0));
      double al = rmax > 0.0 ? eta / rmax : 1.0;
      if (al > 1.0) al = 1.0;
      al_pending = al;
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
