// stub (written for this repo): see lammps_stub.h
#pragma once
#include "lammps_stub.h"
