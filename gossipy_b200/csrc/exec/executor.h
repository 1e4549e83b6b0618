// Native executor of the gossip round loop (see executor.cpp).
#pragma once
#include <pybind11/pybind11.h>

namespace gb {
void bind_executor(pybind11::module_& m);
}  // namespace gb
