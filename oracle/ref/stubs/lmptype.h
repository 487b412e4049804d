#pragma once
#include <cstdint>
namespace LAMMPS_NS { typedef int tagint; typedef int64_t bigint; }
