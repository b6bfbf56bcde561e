/* oracle/mps_oracle.h — C ABI of the CPU restatement of the reference MPS step.  TEST INFRASTRUCTURE ONLY
 * (see the header comment of mps_oracle.cpp).  Mirrors oracle/ref_capi.cpp (ref_* -> orc_*) so that the same Python
 * driver (oracle/refbind.py, oracle/portbind.py) can run either.
 */
#ifndef OPENMPS_B200_MPS_ORACLE_H
#define OPENMPS_B200_MPS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void* orc_create(int dim, int central_gravity, double maxDt, double courant, double g, double rho, double nu,
                 double r_eByl_0, double l_0, const double* minX, const double* maxX, double eps);
void orc_destroy(void* h);
const char* orc_last_error(void* h);
void orc_add_particles(void* h, uint64_t n, const double* x, const double* u, const double* p, const double* nd, const int32_t* type);
void orc_set_wall_positions(void* h, uint64_t n, const uint64_t* ids, const double* x);
uint64_t orc_count(void* h);
void orc_get_state(void* h, double* x, double* u, double* p, double* nd, int32_t* type);
void orc_set_state(void* h, const double* x, const double* u, const double* p, const double* nd);
void orc_get_env(void* h, double* out10);
void orc_set_dt(void* h, double dt, int advance);
double orc_determine_dt(void* h);
int orc_stage(void* h, const char* name);
int orc_forward(void* h, uint64_t steps, double dt, uint64_t* done, double* seconds);
int orc_run_until(void* h, double tEnd, uint64_t* done);
void orc_get_cells(void* h, int64_t* cell);
uint64_t orc_grid_capacity(void* h);
void orc_get_neighbors(void* h, uint64_t* rowptr, uint64_t* idx);
uint64_t orc_csr_nnz(void* h);
void orc_get_csr(void* h, uint32_t* rowptr, uint32_t* col, double* val);
void orc_set_system(void* h, uint64_t n, const uint32_t* rowptr, const uint32_t* col, const double* val, const double* b, const double* x0);
int orc_get_vec(void* h, int which, double* out);
double orc_dndt(void* h, uint64_t i);
uint64_t orc_last_iterations(void* h);
#ifdef __cplusplus
}
#endif
#endif
