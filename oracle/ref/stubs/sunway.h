// stub (written for this repo): Sunway runtime hooks absent on x86
#pragma once
