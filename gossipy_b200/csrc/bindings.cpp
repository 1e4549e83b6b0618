// Python bindings of the gossipy_b200 sm_100a extension (module gossipy_b200._C).
#include "ops.h"

namespace gb {
// runtime.cpp
int64_t ipc_alloc(int64_t nbytes);
void ipc_free(int64_t ptr);
pybind11::bytes ipc_get_handle(int64_t ptr);
int64_t ipc_open_handle(pybind11::bytes handle);
void ipc_close_handle(int64_t ptr);
at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, int64_t device, bool as_int32);
void flag_signal(int64_t flag_ptr, int64_t value);
void flag_wait(int64_t flag_ptr, int64_t value);
int64_t device_sm_count();
}  // namespace gb

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "gossipy_b200 sm_100a kernels";
    m.def("merge_pair", &gb::merge_pair);
    m.def("merge_segments", &gb::merge_segments);
    m.def("merge_indexed", &gb::merge_indexed);
    m.def("merge_kway", &gb::merge_kway);
    m.def("sgd_step", &gb::sgd_step);
    m.def("adam_step", &gb::adam_step);
    m.def("mlp1_train", &gb::mlp1_train);
    m.def("mlp1_train_tc_debug", &gb::mlp1_train_tc_debug);
    m.def("mlp1_eval", &gb::mlp1_eval);
    m.def("logreg_train", &gb::logreg_train);
    m.def("logreg_scores", &gb::logreg_scores);
    m.def("linear_seq_update", &gb::linear_seq_update);
    m.def("kmeans_update", &gb::kmeans_update);
    m.def("kmeans_assign", &gb::kmeans_assign);
    m.def("mf_update", &gb::mf_update);
    m.def("tc_probe", &gb::tc_probe);
    m.def("ipc_alloc", &gb::ipc_alloc);
    m.def("ipc_free", &gb::ipc_free);
    m.def("ipc_get_handle", &gb::ipc_get_handle);
    m.def("ipc_open_handle", &gb::ipc_open_handle);
    m.def("ipc_close_handle", &gb::ipc_close_handle);
    m.def("tensor_from_ptr", &gb::tensor_from_ptr);
    m.def("flag_signal", &gb::flag_signal);
    m.def("flag_wait", &gb::flag_wait);
    m.def("device_sm_count", &gb::device_sm_count);
}
