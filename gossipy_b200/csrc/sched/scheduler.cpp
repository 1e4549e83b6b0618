// Native discrete-time gossip scheduler -- placeholder, filled in below.
#include "scheduler.h"

namespace gb {
void bind_scheduler(pybind11::module_& m) { (void)m; }
}  // namespace gb
