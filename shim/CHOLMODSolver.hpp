// Adapter for the reference's compile-time solver switch (Types.hpp:15, Optimizer.cpp:13-19, 89-95):
// compiling the UNMODIFIED Optimizer.cpp with -DLINSYSSOLVER_USE_CHOLMOD and this directory first on the
// include path makes `new CHOLMODSolver<Eigen::VectorXi, Eigen::VectorXd>()` construct the CUDA solver.
// (The reference's own CHOLMODSolver needs SuiteSparse, which it does not vendor; this header shadows it.)
#ifndef CHOLMODSolver_hpp
#define CHOLMODSolver_hpp
#include "CudaLinSysSolver.hpp"
namespace OptCuts {
template <typename vectorTypeI, typename vectorTypeS>
using CHOLMODSolver = CudaLinSysSolver<vectorTypeI, vectorTypeS>;
}
#endif
