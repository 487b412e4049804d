/* stub (written for this repo): Sunway SIMD intrinsics are only used inside the CPE (slave-core) sections */
#pragma once
