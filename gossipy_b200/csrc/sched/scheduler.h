// Native discrete-time gossip scheduler (see scheduler.cpp).
#pragma once
#include <pybind11/pybind11.h>

namespace gb {
void bind_scheduler(pybind11::module_& m);
}  // namespace gb
