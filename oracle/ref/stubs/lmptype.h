// stub of LAMMPS' lmptype.h (written for this repo): LAMMPS_SMALLBIG integer types
#pragma once
#include <stdint.h>
#define BIGINT_FORMAT "%ld"
#define TAGINT_FORMAT "%d"
#ifdef __cplusplus
namespace LAMMPS_NS { typedef int tagint; typedef int64_t bigint; }
#endif
