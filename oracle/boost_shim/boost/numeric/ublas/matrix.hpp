// Stand-in for <boost/numeric/ublas/matrix.hpp> — TEST INFRASTRUCTURE ONLY (oracle build).
// Computer.hpp:15 includes it; c_matrix/outer_prod/prod are only used under MPS_GC (off in the default build,
// defines.hpp:46), so nothing needs to be provided here.
#ifndef OPENMPS_B200_ORACLE_UBLAS_MATRIX_SHIM
#define OPENMPS_B200_ORACLE_UBLAS_MATRIX_SHIM
#include "vector.hpp"
#endif
