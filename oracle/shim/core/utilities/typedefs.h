// Stand-in for legate.core's core/utilities/typedefs.h (external, not under /root/reference).
// TEST INFRASTRUCTURE ONLY — see oracle/README.md.
#pragma once
#include "legate.h"
