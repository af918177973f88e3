// Stand-in for <gurobi_c++.h> (TEST INFRASTRUCTURE, oracle/_ref build): the part of the Gurobi C++ API that the
// reference's neptune/src/solver_gurobi_poly.cpp and neptune/include/solver_gurobi_utils.hpp use.  The MODEL is recorded
// exactly as the reference builds it (variables, linear rows with sense and right-hand side, quadratic objective,
// quadratic rows); GRBModel::optimize hands it to a solver callback installed by the test (HiGHS through scipy) and keeps
// the status and the primal values it returns.  Nothing of Gurobi's algorithm is restated here.
//
// Lazy-update semantics are reproduced because the reference depends on them: variables and rows added since the last
// update() are not visible to getVars() / getConstrs() / NumVars / NumConstrs, and remove() takes effect at the next
// update() -- optimize() starts with resetCompleteModel(m_) (solver_gurobi_utils.hpp), which must remove the PREVIOUS
// replan's model and leave the variables setInitTrajectory / setHulls have just added (solver_gurobi_poly.cpp:187-281).
#pragma once
#include <cstddef>
#include <limits>
#include <string>
#include <vector>

#define GRB_INFINITY 1e100
#define GRB_CONTINUOUS 'C'
#define GRB_MINIMIZE 1
#define GRB_MAXIMIZE -1
#define GRB_LOADED 1
#define GRB_OPTIMAL 2
#define GRB_INFEASIBLE 3
#define GRB_INF_OR_UNBD 4
#define GRB_UNBOUNDED 5
#define GRB_CUTOFF 6
#define GRB_ITERATION_LIMIT 7
#define GRB_NODE_LIMIT 8
#define GRB_TIME_LIMIT 9
#define GRB_SOLUTION_LIMIT 10
#define GRB_INTERRUPTED 11
#define GRB_NUMERIC 12
#define GRB_SUBOPTIMAL 13
#define GRB_INPROGRESS 14
#define GRB_USER_OBJ_LIMIT 15

enum GRB_IntAttr
{
  GRB_IntAttr_NumConstrs,
  GRB_IntAttr_NumVars,
  GRB_IntAttr_NumQConstrs,
  GRB_IntAttr_NumGenConstrs,
  GRB_IntAttr_Status,
  GRB_IntAttr_SolCount
};

class GRBModel;

// ---- recorded model handed to the solver hook (dense, variables = the model's ACTIVE variables in creation order)
struct ref_qp_model
{
  int nvar;
  const double* lb;      // [nvar]
  const double* ub;      // [nvar]
  const double* Q;       // [nvar][nvar] objective = x'Qx + c'x + c0 (Q as accumulated, not symmetrised)
  const double* c;       // [nvar]
  double c0;
  int nlin;
  const double* A;       // [nlin][nvar]
  const char* sense;     // [nlin] '<', '>', '='
  const double* rhs;     // [nlin]
  int nquad;
  const double* Qc;      // [nquad][nvar][nvar]   rows: x'Qc x + qc'x (sense) qrhs
  const double* qc;      // [nquad][nvar]
  const char* qsense;    // [nquad]
  const double* qrhs;    // [nquad]
  double time_limit;     // "TimeLimit" parameter as set by the reference
  int non_convex;        // "NonConvex" parameter
};
// writes x [nvar], returns a GRB_* status; sol_count = 1 iff x is a solution
typedef int (*ref_qp_solver)(const ref_qp_model* m, double* x, int* sol_count, void* user);
extern "C" void ref_set_qp_solver(ref_qp_solver f, void* user);

struct GRBEnv
{
};

class GRBVar
{
public:
  GRBModel* m = nullptr;
  int id = -1;
};

struct ref_lin_term
{
  int var;
  double coef;
};
struct ref_quad_term
{
  int v1, v2;
  double coef;
};

class GRBLinExpr
{
public:
  GRBModel* m = nullptr;
  double cst = 0.0;
  std::vector<ref_lin_term> t;
  GRBLinExpr(double c = 0.0) : cst(c) {}
  GRBLinExpr(GRBVar v) : m(v.m) { t.push_back({ v.id, 1.0 }); }
  GRBLinExpr& operator+=(const GRBLinExpr& o)
  {
    if (!m) m = o.m;
    cst += o.cst;
    t.insert(t.end(), o.t.begin(), o.t.end());
    return *this;
  }
  GRBLinExpr& operator-=(const GRBLinExpr& o)
  {
    if (!m) m = o.m;
    cst -= o.cst;
    for (auto& q : o.t) t.push_back({ q.var, -q.coef });
    return *this;
  }
  GRBLinExpr& operator*=(double s)
  {
    cst *= s;
    for (auto& q : t) q.coef *= s;
    return *this;
  }
  double getValue() const;
};
inline GRBLinExpr operator+(GRBLinExpr a, const GRBLinExpr& b) { return a += b; }
inline GRBLinExpr operator-(GRBLinExpr a, const GRBLinExpr& b) { return a -= b; }
inline GRBLinExpr operator-(GRBLinExpr a) { return a *= -1.0; }
inline GRBLinExpr operator*(double s, GRBLinExpr a) { return a *= s; }
inline GRBLinExpr operator*(GRBLinExpr a, double s) { return a *= s; }

class GRBQuadExpr
{
public:
  GRBLinExpr lin;
  std::vector<ref_quad_term> q;
  GRBQuadExpr(double c = 0.0) : lin(c) {}
  GRBQuadExpr(const GRBLinExpr& l) : lin(l) {}
  GRBQuadExpr& operator+=(const GRBQuadExpr& o)
  {
    lin += o.lin;
    q.insert(q.end(), o.q.begin(), o.q.end());
    return *this;
  }
  GRBQuadExpr& operator-=(const GRBQuadExpr& o)
  {
    lin -= o.lin;
    for (auto& e : o.q) q.push_back({ e.v1, e.v2, -e.coef });
    return *this;
  }
  GRBQuadExpr& operator*=(double s)
  {
    lin *= s;
    for (auto& e : q) e.coef *= s;
    return *this;
  }
  double getValue() const;
};
inline GRBQuadExpr operator*(const GRBLinExpr& a, const GRBLinExpr& b)
{
  GRBQuadExpr r(a.cst * b.cst);
  r.lin.m = a.m ? a.m : b.m;
  for (auto& x : a.t) r.lin.t.push_back({ x.var, x.coef * b.cst });
  for (auto& y : b.t) r.lin.t.push_back({ y.var, y.coef * a.cst });
  for (auto& x : a.t)
    for (auto& y : b.t) r.q.push_back({ x.var, y.var, x.coef * y.coef });
  return r;
}
inline GRBQuadExpr operator+(GRBQuadExpr a, const GRBQuadExpr& b) { return a += b; }
inline GRBQuadExpr operator-(GRBQuadExpr a, const GRBQuadExpr& b) { return a -= b; }
inline GRBQuadExpr operator*(double s, GRBQuadExpr a) { return a *= s; }
inline GRBQuadExpr operator*(GRBQuadExpr a, double s) { return a *= s; }

class GRBTempConstr
{
public:
  GRBQuadExpr e;  // e (sense) 0
  char sense;
};
inline GRBTempConstr operator<=(const GRBLinExpr& a, const GRBLinExpr& b) { return { GRBQuadExpr(a - b), '<' }; }
inline GRBTempConstr operator>=(const GRBLinExpr& a, const GRBLinExpr& b) { return { GRBQuadExpr(a - b), '>' }; }
inline GRBTempConstr operator==(const GRBLinExpr& a, const GRBLinExpr& b) { return { GRBQuadExpr(a - b), '=' }; }
inline GRBTempConstr operator<=(const GRBQuadExpr& a, const GRBQuadExpr& b) { return { a - b, '<' }; }
inline GRBTempConstr operator>=(const GRBQuadExpr& a, const GRBQuadExpr& b) { return { a - b, '>' }; }
inline GRBTempConstr operator==(const GRBQuadExpr& a, const GRBQuadExpr& b) { return { a - b, '=' }; }

class GRBConstr
{
public:
  int id = -1;
};
class GRBQConstr
{
public:
  int id = -1;
};
class GRBGenConstr
{
public:
  int id = -1;
};

class GRBModel
{
public:
  enum State
  {
    PENDING = 0,  // added since the last update()
    ACTIVE = 1,
    REMOVED = 2   // removed (takes effect at update(): then it is dropped from every listing)
  };
  struct Var
  {
    double lb, ub;
    std::string name;
    int state;
    bool to_remove;
  };
  struct Row
  {
    GRBQuadExpr e;
    char sense;
    int state;
    bool to_remove;
  };
  explicit GRBModel(const GRBEnv&) {}
  GRBVar addVar(double lb, double ub, double obj, char, std::string name)
  {
    (void)obj;
    vars.push_back({ lb, ub, name, PENDING, false });
    GRBVar v;
    v.m = this;
    v.id = (int)vars.size() - 1;
    return v;
  }
  GRBConstr addConstr(const GRBTempConstr& c)
  {
    lin.push_back({ c.e, c.sense, PENDING, false });
    GRBConstr r;
    r.id = (int)lin.size() - 1;
    return r;
  }
  GRBQConstr addQConstr(const GRBTempConstr& c)
  {
    quad.push_back({ c.e, c.sense, PENDING, false });
    GRBQConstr r;
    r.id = (int)quad.size() - 1;
    return r;
  }
  void setObjective(const GRBQuadExpr& e, int sense)
  {
    objective = e;
    obj_sense = sense;
  }
  GRBQuadExpr getObjective() const
  {
    GRBQuadExpr o = objective;
    o.lin.m = const_cast<GRBModel*>(this);
    return o;
  }
  void update()
  {
    for (auto& v : vars)
    {
      if (v.to_remove) v.state = REMOVED;
      if (v.state == PENDING) v.state = ACTIVE;
    }
    for (auto* rows : { &lin, &quad })
      for (auto& r : *rows)
      {
        if (r.to_remove) r.state = REMOVED;
        if (r.state == PENDING) r.state = ACTIVE;
      }
  }
  void reset() { x.clear(), status = GRB_LOADED, sol_count = 0; }
  void set(const std::string& name, const std::string& value)
  {
    if (name == "TimeLimit") time_limit = std::stod(value);
    if (name == "NonConvex") non_convex = std::stoi(value);
  }
  int get(GRB_IntAttr a) const
  {
    switch (a)
    {
      case GRB_IntAttr_NumVars:
      {
        int n = 0;
        for (auto& v : vars) n += v.state == ACTIVE;
        return n;
      }
      case GRB_IntAttr_NumConstrs:
      {
        int n = 0;
        for (auto& r : lin) n += r.state == ACTIVE;
        return n;
      }
      case GRB_IntAttr_NumQConstrs:
      {
        int n = 0;
        for (auto& r : quad) n += r.state == ACTIVE;
        return n;
      }
      case GRB_IntAttr_NumGenConstrs:
        return 0;
      case GRB_IntAttr_Status:
        return status;
      case GRB_IntAttr_SolCount:
        return sol_count;
    }
    return 0;
  }
  // arrays of the ACTIVE objects, heap-allocated like Gurobi's (the reference leaks them, so does this)
  GRBVar* getVars()
  {
    GRBVar* out = new GRBVar[get(GRB_IntAttr_NumVars) + 1];
    int n = 0;
    for (size_t i = 0; i < vars.size(); i++)
      if (vars[i].state == ACTIVE) out[n].m = this, out[n++].id = (int)i;
    return out;
  }
  GRBConstr* getConstrs()
  {
    GRBConstr* out = new GRBConstr[get(GRB_IntAttr_NumConstrs) + 1];
    int n = 0;
    for (size_t i = 0; i < lin.size(); i++)
      if (lin[i].state == ACTIVE) out[n++].id = (int)i;
    return out;
  }
  GRBQConstr* getQConstrs()
  {
    GRBQConstr* out = new GRBQConstr[get(GRB_IntAttr_NumQConstrs) + 1];
    int n = 0;
    for (size_t i = 0; i < quad.size(); i++)
      if (quad[i].state == ACTIVE) out[n++].id = (int)i;
    return out;
  }
  GRBGenConstr* getGenConstrs() { return new GRBGenConstr[1]; }
  void remove(GRBVar v) { vars[v.id].to_remove = true; }
  void remove(GRBConstr c) { lin[c.id].to_remove = true; }
  void remove(GRBQConstr c) { quad[c.id].to_remove = true; }
  void remove(GRBGenConstr) {}
  void optimize();  // ref_gurobi_capture.cpp

  std::vector<Var> vars;
  std::vector<Row> lin, quad;
  GRBQuadExpr objective;
  int obj_sense = GRB_MINIMIZE;
  std::vector<double> x;  // value of every variable ever created (by id) after optimize()
  int status = GRB_LOADED, sol_count = 0;
  double time_limit = 1e100;
  int non_convex = 0;
};

inline double GRBLinExpr::getValue() const
{
  double v = cst;
  for (auto& q : t) v += q.coef * (m && q.var < (int)m->x.size() ? m->x[q.var] : 0.0);
  return v;
}
inline double GRBQuadExpr::getValue() const
{
  double v = lin.getValue();
  const GRBModel* m = lin.m;
  for (auto& e : q)
    if (m && e.v1 < (int)m->x.size() && e.v2 < (int)m->x.size()) v += e.coef * m->x[e.v1] * m->x[e.v2];
  return v;
}
