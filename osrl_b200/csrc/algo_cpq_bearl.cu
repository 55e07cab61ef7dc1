// Step programs of CPQ (cpq.py:125-230) and BEAR-Lag (bearl.py:144-335).
#include "engine.h"

namespace osrl {
void build_cpq(Engine&) { throw Err(OSRL_ERR_UNSUPPORTED, "CPQ step program not built yet"); }
void build_bearl(Engine&) { throw Err(OSRL_ERR_UNSUPPORTED, "BEAR-Lag step program not built yet"); }
}  // namespace osrl
