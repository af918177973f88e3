// nb_tables.cpp -- host-side precomputation of the reduced-space QP tables (NbQpTable) and the
// solver constants (NbConsts).  Runs once in nb_create.
//
// The reference rebuilds a 12n-variable Gurobi model every replan
// (solver_gurobi_poly.cpp:187-224, :385-425, :659-678).  Its equality rows depend only on
// (n, T_span) and, linearly, on the initial (b0, c0, d0) of each axis, so they are eliminated here
// once: x_ax = Pm * init3 + Z * w.  Everything the interior-point kernel needs per inequality row
// ("features": MINVO position / velocity control points and end accelerations, :441-470) becomes a
// precomputed dof-vector.
#include "nb_tables.h"

#include <string.h>

#include <cmath>
#include <vector>

namespace
{
// MINVO matrices for t in [0,1]: reference neptune/include/mader_types.hpp:152-162
const double kAposMv[16] = {
  -3.4416308968564117698463178385282, 6.9895481477801393310755884158425, -4.4622887507045296828778191411402,
  0.91437149978080234369315348885721, 6.6792587327074839365081970754545, -11.845989901556746914934592496138,
  5.2523596690684613008670567069203, 0.0, -6.6792587327074839365081970754545, 8.1917862965657040064115790301003,
  -1.5981560640774179482548333908198, 0.085628500219197656306846511142794, 3.4416308968564117698463178385282,
  -3.3353445427890959784633650997421, 0.80808514571348655231020075007109,
  -0.0000000000000000084567769453869345852581318467855 };
const double kAvelMv[9] = { 1.5, -2.36602540378444, 0.933012701892219, -3.0, 3.0, 0.0,
                            1.5, -0.633974596215561, 0.0669872981077807 };

bool invert(const double* A, int n, double* out)
{
  std::vector<double> w(n * 2 * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      w[i * 2 * n + j] = A[i * n + j];
      w[i * 2 * n + n + j] = i == j ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; c++)
  {
    int p = c;
    for (int r = c + 1; r < n; r++)
      if (std::fabs(w[r * 2 * n + c]) > std::fabs(w[p * 2 * n + c])) p = r;
    if (w[p * 2 * n + c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 2 * n; j++) std::swap(w[c * 2 * n + j], w[p * 2 * n + j]);
    const double piv = 1.0 / w[c * 2 * n + c];
    for (int j = 0; j < 2 * n; j++) w[c * 2 * n + j] *= piv;
    for (int r = 0; r < n; r++)
      if (r != c && w[r * 2 * n + c] != 0.0)
      {
        const double f = w[r * 2 * n + c];
        for (int j = 0; j < 2 * n; j++) w[r * 2 * n + j] -= f * w[c * 2 * n + j];
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) out[i * n + j] = w[i * 2 * n + n + j];
  return true;
}

// modified Gram-Schmidt (twice) on the columns of B (rows x cols, row-major); returns rank kept
int orthonormalize(std::vector<double>& B, int rows, int cols)
{
  int kept = 0;
  for (int c = 0; c < cols; c++)
  {
    std::vector<double> v(rows);
    double n0 = 0;
    for (int r = 0; r < rows; r++)
    {
      v[r] = B[r * cols + c];
      n0 += v[r] * v[r];
    }
    n0 = std::sqrt(n0);
    for (int pass = 0; pass < 2; pass++)
      for (int k = 0; k < kept; k++)
      {
        double d = 0;
        for (int r = 0; r < rows; r++) d += v[r] * B[r * cols + k];
        for (int r = 0; r < rows; r++) v[r] -= d * B[r * cols + k];
      }
    double nn = 0;
    for (int r = 0; r < rows; r++) nn += v[r] * v[r];
    nn = std::sqrt(nn);
    if (nn > 1e-11 * (n0 > 0 ? n0 : 1.0))
    {
      for (int r = 0; r < rows; r++) B[r * cols + kept] = v[r] / nn;
      kept++;
    }
  }
  return kept;
}
}  // namespace

void nb_build_consts(const nb_params* p, NbConsts* c)
{
  memset(c, 0, sizeof(*c));
  const double T = p->T_span;
  c->T = T;
  c->W = p->weight;
  double A[16], Av[9];
  const double cp[4] = { 1.0 / (T * T * T), 1.0 / (T * T), 1.0 / T, 1.0 };
  const double cv[3] = { 1.0 / (T * T), 1.0 / T, 1.0 };
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) A[i * 4 + j] = kAposMv[i * 4 + j] * cp[j];  // :61
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Av[i * 3 + j] = kAvelMv[i * 3 + j] * cv[j];  // :62
  invert(A, 4, c->Ainv);                                                     // :93
  invert(Av, 3, c->V);                                                       // :94
  for (int j = 0; j < 3; j++)
  {
    c->V[j] *= 3.0;      // :96
    c->V[3 + j] *= 2.0;  // :97
  }
  invert(kAposMv, 4, c->Ainv01);
  for (int k = 0; k < 3; k++)
  {
    c->lim_min[k] = p->lim_min[k];
    c->lim_max[k] = p->lim_max[k];
  }
  c->v_max = p->v_max;
  c->a_max = p->a_max;
  const double dx = p->lim_max[0] - p->lim_min[0], dy = p->lim_max[1] - p->lim_min[1];
  c->long_length = std::sqrt(dx * dx + dy * dy);  // :173
  c->drone_radius = p->drone_radius;
  c->N = p->num_agents;
  c->M = p->num_static;
  c->num_pol = p->num_pol;
  c->S = p->samples;
  c->ent_cap = p->ent_cap;
  c->bp_max = p->bp_max;
  c->ent_slots = p->ent_slots;
  c->max_iter = p->ipm_max_iter;
  c->tol = p->ipm_tol;
}

bool nb_build_table(const NbConsts* cs, int n, int mode, NbQpTable* t)
{
  memset(t, 0, sizeof(*t));
  t->n = n;
  t->mode = mode;
  const double T = cs->T, W = cs->W;
  const double qp[4] = { T * T * T, T * T, T, 1.0 };      // :126
  const double qv[4] = { 3 * T * T, 2 * T, 1.0, 0.0 };    // :128
  const double qa[4] = { 6 * T, 2.0, 0.0, 0.0 };          // :129
  const int R = 4 * n;
  // propagation: coefficients as a linear function of (a_0..a_{n-1}) [Za] and of (b0,c0,d0) [Xp0],
  // using E1 (:390-396) and the continuity rows E2 (:400-425)
  std::vector<double> Za(R * n, 0.0), Xp0(R * 3, 0.0);
  auto propagate = [&](std::vector<double>& X, int cols) {
    for (int i = 0; i + 1 < n; i++)
      for (int c = 0; c < cols; c++)
      {
        double p = 0, v = 0, a = 0;
        for (int k = 0; k < 4; k++)
        {
          const double x = X[(4 * i + k) * cols + c];
          p += qp[k] * x;
          v += qv[k] * x;
          a += qa[k] * x;
        }
        X[(4 * (i + 1) + 3) * cols + c] += p;
        X[(4 * (i + 1) + 2) * cols + c] += v;
        X[(4 * (i + 1) + 1) * cols + c] += a / 2.0;
      }
  };
  for (int j = 0; j < n; j++) Za[(4 * j + 0) * n + j] = 1.0;
  Xp0[1 * 3 + 0] = 1.0;  // b0
  Xp0[2 * 3 + 1] = 1.0;  // c0
  Xp0[3 * 3 + 2] = 1.0;  // d0
  propagate(Za, n);
  propagate(Xp0, 3);

  std::vector<double> Bz;  // basis of the free space before orthonormalisation (R x cols)
  std::vector<double> Pm(R * 3, 0.0);
  int cols = 0;
  if (mode == 1)
  {
    Bz = Za;
    cols = n;
    Pm = Xp0;
  }
  else
  {
    // terminal v/a rows E3 (:660-678): Tm a = -tr0 init3
    std::vector<double> Tm(2 * n, 0.0);
    double tr0[6] = { 0, 0, 0, 0, 0, 0 };
    for (int k = 0; k < 4; k++)
    {
      for (int j = 0; j < n; j++)
      {
        Tm[0 * n + j] += qv[k] * Za[(4 * (n - 1) + k) * n + j];
        Tm[1 * n + j] += qa[k] * Za[(4 * (n - 1) + k) * n + j];
      }
      for (int c = 0; c < 3; c++)
      {
        tr0[0 * 3 + c] += qv[k] * Xp0[(4 * (n - 1) + k) * 3 + c];
        tr0[1 * 3 + c] += qa[k] * Xp0[(4 * (n - 1) + k) * 3 + c];
      }
    }
    std::vector<double> apmap(n * 3, 0.0);  // a_p = apmap * init3
    if (n >= 2)
    {
      double G2[4] = { 0, 0, 0, 0 }, G2i[4];
      for (int j = 0; j < n; j++)
      {
        G2[0] += Tm[j] * Tm[j];
        G2[1] += Tm[j] * Tm[n + j];
        G2[3] += Tm[n + j] * Tm[n + j];
      }
      G2[2] = G2[1];
      if (!invert(G2, 2, G2i)) return false;
      // a_p = Tm^T (Tm Tm^T)^-1 (-tr0 init3)
      for (int j = 0; j < n; j++)
        for (int c = 0; c < 3; c++)
        {
          double y0 = -(G2i[0] * tr0[c] + G2i[1] * tr0[3 + c]);
          double y1 = -(G2i[2] * tr0[c] + G2i[3] * tr0[3 + c]);
          apmap[j * 3 + c] = Tm[j] * y0 + Tm[n + j] * y1;
        }
      // null space of Tm: orthonormalise [Tm^T | I] and drop the first two directions
      std::vector<double> Q(n * (n + 2), 0.0);
      for (int j = 0; j < n; j++)
      {
        Q[j * (n + 2) + 0] = Tm[j];
        Q[j * (n + 2) + 1] = Tm[n + j];
        Q[j * (n + 2) + 2 + j] = 1.0;
      }
      const int kept = orthonormalize(Q, n, n + 2);
      if (kept != n) return false;
      cols = n - 2;
      Bz.assign(R * (cols > 0 ? cols : 1), 0.0);
      for (int r = 0; r < R; r++)
        for (int c = 0; c < cols; c++)
        {
          double s = 0;
          for (int j = 0; j < n; j++) s += Za[r * n + j] * Q[j * (n + 2) + 2 + c];
          Bz[r * cols + c] = s;
        }
    }
    else
    {
      // n = 1: two rows, one unknown -> least squares a = Tm^+ r, and a consistency residual
      const double tt = Tm[0] * Tm[0] + Tm[1] * Tm[1];
      for (int c = 0; c < 3; c++)
      {
        const double r0 = -tr0[c], r1 = -tr0[3 + c];
        const double a = (Tm[0] * r0 + Tm[1] * r1) / tt;
        apmap[c] = a;
        t->Rres[0][c] = r0 - Tm[0] * a;
        t->Rres[1][c] = r1 - Tm[1] * a;
      }
      t->has_resid = 1;
      cols = 0;
    }
    for (int r = 0; r < R; r++)
      for (int c = 0; c < 3; c++)
      {
        double s = Xp0[r * 3 + c];
        for (int j = 0; j < n; j++) s += Za[r * n + j] * apmap[j * 3 + c];
        Pm[r * 3 + c] = s;
      }
  }
  int dof = 0;
  if (cols > 0)
  {
    dof = orthonormalize(Bz, R, cols);
    if (dof != cols) return false;
  }
  t->dof = dof;
  for (int r = 0; r < R; r++)
  {
    for (int c = 0; c < dof; c++) t->Z[r][c] = Bz[r * cols + c];
    for (int c = 0; c < 3; c++) t->Pm[r][c] = Pm[r * 3 + c];
  }
  // feature rows (:441-470): j = 0..3 MINVO position CP k, 4..6 velocity CP k, 7 end acceleration
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 8; j++)
    {
      double row[4] = { 0, 0, 0, 0 };
      if (j < 4)
        for (int q = 0; q < 4; q++) row[q] = cs->Ainv[q * 4 + j];
      else if (j < 7)
        for (int q = 0; q < 3; q++) row[q] = cs->V[q * 3 + (j - 4)];
      else
      {
        row[0] = 6.0 * T;
        row[1] = 2.0;
      }
      for (int q = 0; q < 4; q++)
      {
        for (int c = 0; c < dof; c++) t->C[i * 8 + j][c] += row[q] * t->Z[4 * i + q][c];
        for (int c = 0; c < 3; c++) t->c0[i * 8 + j][c] += row[q] * t->Pm[4 * i + q][c];
      }
    }
  for (int f = 0; f < 8 * n; f++)
  {
    for (int c = 0; c < dof; c++) t->Ct[c][f] = t->C[f][c];
    for (int ca = 0; ca < dof; ca++)
      for (int cb = 0; cb <= ca; cb++) t->PP[f][ca * (ca + 1) / 2 + cb] = t->C[f][ca] * t->C[f][cb];
  }
  // objective (:322-380): 36 T sum a_i^2 + W (qp.x_last - pf)^2 [+ W ((qv.x_last)^2 + (qa.x_last)^2)]
  double tv[NB_DOF_MAX] = { 0 }, ta[NB_DOF_MAX] = { 0 }, tv0[3] = { 0 }, ta0[3] = { 0 };
  for (int q = 0; q < 4; q++)
  {
    for (int c = 0; c < dof; c++)
    {
      t->tq[c] += qp[q] * t->Z[4 * (n - 1) + q][c];
      tv[c] += qv[q] * t->Z[4 * (n - 1) + q][c];
      ta[c] += qa[q] * t->Z[4 * (n - 1) + q][c];
    }
    for (int c = 0; c < 3; c++)
    {
      t->tq0[c] += qp[q] * t->Pm[4 * (n - 1) + q][c];
      tv0[c] += qv[q] * t->Pm[4 * (n - 1) + q][c];
      ta0[c] += qa[q] * t->Pm[4 * (n - 1) + q][c];
    }
  }
  for (int a = 0; a < dof; a++)
  {
    for (int b = 0; b < dof; b++)
    {
      double h = 0;
      for (int i = 0; i < n; i++) h += 72.0 * T * t->Z[4 * i][a] * t->Z[4 * i][b];
      h += 2.0 * W * t->tq[a] * t->tq[b];
      if (mode == 1) h += 2.0 * W * (tv[a] * tv[b] + ta[a] * ta[b]);
      t->Hr[a][b] = h;
    }
    for (int c = 0; c < 3; c++)
    {
      double g = 0;
      for (int i = 0; i < n; i++) g += 72.0 * T * t->Z[4 * i][a] * t->Pm[4 * i][c];
      g += 2.0 * W * t->tq[a] * t->tq0[c];
      if (mode == 1) g += 2.0 * W * (tv[a] * tv0[c] + ta[a] * ta0[c]);
      t->Gr[a][c] = g;
    }
    t->gpf[a] = -2.0 * W * t->tq[a];
  }
  return true;
}
