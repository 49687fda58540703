/* TEST INFRASTRUCTURE — plain-C restatement of the OptCuts hot path (oracle).  See sd_port.c. */
#ifndef SD_PORT_H
#define SD_PORT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* all matrices column-major as Eigen stores them; vectors over DOFs interleaved [u0 v0 u1 v1 ...] */
int    port_rest_features(int nV, int nF, const double* V_rest, const int32_t* F, double areaThres_AM,
                          double* rest8, double* scalars3);
void   port_energy_per_elem(int nV, int nF, const int32_t* F, const double* UV, const double* rest8,
                            double surfaceArea, int uniform, double* out);
double port_energy(int nV, int nF, const int32_t* F, const double* UV, const double* rest8,
                   double surfaceArea, int uniform);
void   port_gradient(int nV, int nF, const int32_t* F, const double* UV, const double* rest8,
                     double surfaceArea, int uniform, const int32_t* fixed, int nFixed, double* g);
void   port_make_pd6(double* M36);
void   port_hessian_blocks(int nV, int nF, const int32_t* F, const double* UV, const double* rest8,
                           double surfaceArea, int uniform, int project, double* out36);
int64_t port_hessian_triplets(int nV, int nF, const int32_t* F, const double* UV, const double* rest8,
                              double surfaceArea, int uniform, const int32_t* fixed, int nFixed,
                              double* V, int32_t* I, int32_t* J);
double port_init_step_size(int nV, int nF, const int32_t* F, const double* UV, const double* searchDir, double stepSize);
int64_t port_set_pattern(int nV, const int32_t* adjPtr, const int32_t* adjIdx, const int32_t* fixed, int nFixed,
                         int32_t* ia, int32_t* ja);
int    port_update_a(int n, const int32_t* ia, const int32_t* ja, int64_t nT, const int32_t* I, const int32_t* J,
                     const double* S, double* a);
int    port_ldlt_solve(int n, const int32_t* ia, const int32_t* ja, const double* a, const double* rhs, double* x);
double port_seam_sparsity(int nV, int nCoh, const int32_t* cohE, const int32_t* boundaryEdge, const double* edgeLen,
                          const double* UV, double avgEdgeLen, double initSeamLen, int triSoup);
void   port_divgrad(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surfaceArea, double* out);
/* one geometry step of Optimizer::solve(1) on mesh (+ optional air mesh); returns 1 if converged/stopped */
typedef struct { double sqn_g, alpha, E_new, E_scaf_new, E_sd_new, lastEDec, E_last; int converged, stopped, n_halvings; } port_newton_result;
int    port_newton_step(int nV, int nF, const int32_t* F, double* UV, const double* rest8, double surfaceArea,
                        const int32_t* fixed, int nFixed,
                        int nVa, int nFa, const int32_t* Fa, double* UVa, const double* rest8a, const int32_t* l2g, int nBnd,
                        const int32_t* fixedAir, int nFixedAir,
                        double energyParam0, double w_scaf, double targetGRes, int allowEDecRelTol,
                        double* searchDir_out, port_newton_result* out);
#ifdef __cplusplus
}
#endif
#endif
