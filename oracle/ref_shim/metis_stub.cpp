#include "metis.h"
extern "C" {
int METIS_SetDefaultOptions(idx_t*) { return METIS_ERROR; }
int METIS_PartGraphKway(idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, real_t*, real_t*, idx_t*, idx_t*, idx_t*) { return METIS_ERROR; }
}
