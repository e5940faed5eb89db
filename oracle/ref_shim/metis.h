// stand-in for METIS (see README.md): declarations patcher.cu names; the stub entry points fail
#pragma once
#include <stdint.h>
typedef int32_t idx_t;
typedef float   real_t;
enum { METIS_NOPTIONS = 40, METIS_OPTION_PTYPE = 0, METIS_OPTION_OBJTYPE = 1, METIS_OPTION_NUMBERING = 17,
       METIS_OPTION_CONTIG = 13, METIS_OPTION_COMPRESS = 14, METIS_OPTION_DBGLVL = 5, METIS_PTYPE_KWAY = 1,
       METIS_OBJTYPE_VOL = 1, METIS_DBG_TIME = 2, METIS_OK = 1, METIS_ERROR_INPUT = -2, METIS_ERROR_MEMORY = -3,
       METIS_ERROR = -4 };
extern "C" {
int METIS_SetDefaultOptions(idx_t*);
int METIS_PartGraphKway(idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, idx_t*, real_t*, real_t*, idx_t*, idx_t*, idx_t*);
}
