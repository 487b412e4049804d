// stub (written for this repo): software-cache macros of the Sunway CPEs are not used by the serial code paths
#pragma once
