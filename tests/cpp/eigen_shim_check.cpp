// Self-test of the Eigen stand-in (oracle/eigen_shim/Eigen/Dense): prints the results of the operations the reference
// sources rely on, one labelled line each, for tests/test_reference_pin.py to compare with numpy.  The semantics checked
// are Eigen's documented ones: comma initialisation fills row by row, storage is column-major, row / col / block / head /
// tail / segment are views, assigning a view copies coefficients (also across vector orientation), a 1 x 1 product reads
// as a scalar, inverse() of a small fixed matrix.
#include <cstdio>
#include <Eigen/Dense>

template <class M>
static void show(const char* name, const M& m)
{
  std::printf("%s", name);
  for (int r = 0; r < m.rows(); r++)
    for (int c = 0; c < m.cols(); c++) std::printf(" %.17g", (double)m(r, c));
  std::printf("\n");
}

int main()
{
  Eigen::Matrix<double, 4, 4> A;
  A << 4, -2, 1, 0.5, 3, 6, -4, 2, 2, 1, 8, -1, 0.25, -3, 2, 5;
  Eigen::Matrix<double, 3, 3> V;
  V << 2, 0.5, -1, 1, 3, 0.25, -2, 1, 4;
  Eigen::Matrix<double, 2, 4> P;
  P << 1, 2, 3, 4, -1, 0.5, 2, -3;
  show("A", A);
  std::printf("A_data");
  for (int i = 0; i < 16; i++) std::printf(" %.17g", A.data()[i]);
  std::printf("\n");
  show("A_inv", A.inverse());
  show("V_inv", V.inverse());
  show("P_A", P * A);
  show("P_blk_V", P.block<2, 3>(0, 0) * V);
  show("A_T", A.transpose());
  Eigen::Matrix<double, 4, 1> t;
  t << 0.125, 0.25, 0.5, 1;
  Eigen::Vector2d pt = P * t;
  show("P_t", pt);
  double s = A.col(1).transpose() * t;  // 1 x 1 -> scalar
  std::printf("colT_t %.17g\n", s);
  Eigen::Matrix<double, 6, 1> e;
  e << 1, 2, 3, 4, 5, 6;
  e.head<2>() = e.head<2>() + e.segment<2>(2) * 0.5 + e.tail<2>() * 0.25;
  e.tail<2>() = e.tail<2>() - e.head<2>();
  show("e", e);
  Eigen::Matrix<double, 2, 4> Q;
  Q.setZero();
  Q.row(0) = t;              // a column vector assigned to a row
  Q.row(1) = 3 * Q.row(0);
  Q.col(3) = Q.col(0);       // view = view copies coefficients
  show("Q", Q);
  Eigen::Matrix<double, 2, Eigen::Dynamic> D(2, 3);
  D << 1, 2, 3, 4, 5, 6;
  show("D_last", D.rightCols(1));
  show("D_mean", D.rowwise().mean());
  Eigen::Vector2d a(3, 4), b(0, 1);
  std::printf("norm %.17g dot %.17g\n", (a - b).norm(), a.dot(b));
  Eigen::Vector3d g(1, 2, 3);
  Eigen::Vector2d g2 = g.head(2);
  show("g2", g2);
  std::printf("abs %.17g\n", (double)abs(-0.75));  // the floating overload, as with real Eigen's includes
  return 0;
}
