// Stand-in for the reference's math/LinearAlgebra.hpp (Tpetra/Belos/Ifpack2 are absent).  Same class and method names
// as LinearAlgebra.hpp:73-119; the sparsity pattern, the global numbering layer*G + global_id and zeroSystem follow
// LinearAlgebra.cpp:31-152,199-206.  Values outside the static pattern abort (Tpetra would drop them silently).
// Solve() hands the assembled CSR system to a callback the test harness registers (a sparse direct solve), i.e. the
// mathematical contract of Belos GMRES at tolerance 1e-8 — the Krylov iteration itself is not reproduced.
#pragma once
#include "triangulation.hpp"
#include <stdexcept>
#include <string>
#include <vector>

namespace Belos { class StatusTestError : public std::logic_error { using std::logic_error::logic_error; }; }

namespace math { namespace LinearAlgebra {
struct SolveConverge { int numIters; double residual; };
typedef long long global_ordinal_type;

class NearestNeighborProblem {
public:
    mesh& m_domain;
    int m_nLayer;
    std::size_t m_ntri, m_nglobal;
    std::vector<int> rowptr;                   // local rows: layer*ntri + local_id
    std::vector<global_ordinal_type> colgid;   // global column ids
    std::vector<double> values, rhs, solution;

    NearestNeighborProblem(mesh& domain, int nLayer = 1);
    ~NearestNeighborProblem();
    void zeroSystem();
    void matrixReplaceGlobalValues(global_ordinal_type row, global_ordinal_type col, double val);
    void matrixSumIntoGlobalValues(global_ordinal_type row, global_ordinal_type col, double val);
    void matrixResumeFill() {}
    void matrixFillComplete() {}
    void rhsSumIntoGlobalValue(global_ordinal_type idx, double val);
    SolveConverge Solve();
    double getSolutionMax();
    double getRhsMax();
    const double* getSolutionView() { return solution.data(); }
    void writeSystemMatrixMarket(std::string) {}
    void writeSolutionMatrixMarket(std::string) {}
private:
    std::size_t local_row(global_ordinal_type g) const;
    double* entry(global_ordinal_type row, global_ordinal_type col);
};
}}
