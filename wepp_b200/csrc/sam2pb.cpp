// placeholder, replaced below
#include "pipeline.h"
#include <cstdio>
namespace wepp { int sam2pb(const Dataset&) { std::fprintf(stderr, "sam2PB: not built yet\n"); return 1; } }
