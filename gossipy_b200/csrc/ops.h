// Host entry points of the gossipy_b200 sm_100a extension.
#pragma once
#include <torch/extension.h>
#include <vector>

namespace gb {

// merge.cu
void merge_pair(at::Tensor dst, at::Tensor src, double w_dst, double w_src, int64_t lo, int64_t hi);
void merge_segments(at::Tensor dst, at::Tensor src, at::Tensor seg, double w_dst, double w_src);
void merge_indexed(at::Tensor dst, at::Tensor src, at::Tensor idx, double w_dst, double w_src);
void merge_kway(at::Tensor dst, std::vector<at::Tensor> srcs, std::vector<double> weights);

// optim.cu
void sgd_step(at::Tensor p, at::Tensor g, int64_t n, double lr, double wd, double momentum,
              c10::optional<at::Tensor> buf, double dampening, bool nesterov, bool first,
              c10::optional<at::Tensor> scale);
void adam_step(at::Tensor p, at::Tensor g, int64_t n, at::Tensor m, at::Tensor v, int64_t step,
               double lr, double beta1, double beta2, double eps, double wd, bool decoupled);

// mlp1_train.cu / mlp1_eval.cu
int64_t mlp1_train(at::Tensor row, at::Tensor X, at::Tensor y, std::tuple<int64_t, int64_t, int64_t> dims,
                   int64_t batch_size, int64_t local_epochs, double lr, double wd, int64_t key,
                   c10::optional<at::Tensor> part_id, c10::optional<at::Tensor> ages,
                   std::string impl);
at::Tensor mlp1_train_tc_debug(at::Tensor row, at::Tensor X, at::Tensor y,
                               std::tuple<int64_t, int64_t, int64_t> dims, int64_t batch_size,
                               int64_t local_epochs, double lr, double wd, int64_t key);
at::Tensor mlp1_eval(at::Tensor row, at::Tensor X, at::Tensor y, std::tuple<int64_t, int64_t, int64_t> dims,
                     int64_t n_classes, c10::optional<at::Tensor> X_lp);

// small.cu
int64_t logreg_train(at::Tensor row, at::Tensor X, at::Tensor y, std::tuple<int64_t, int64_t> dims,
                     int64_t batch_size, int64_t local_epochs, double lr, double wd, int64_t key,
                     c10::optional<at::Tensor> part_id, c10::optional<at::Tensor> ages);
at::Tensor logreg_scores(at::Tensor row, at::Tensor X, std::tuple<int64_t, int64_t> dims);
void linear_seq_update(at::Tensor w, at::Tensor X, at::Tensor y, int64_t kind, double lr, int64_t n_updates);
void kmeans_update(at::Tensor C, at::Tensor X, double alpha);
at::Tensor kmeans_assign(at::Tensor C, at::Tensor X);
void mf_update(at::Tensor X, at::Tensor b, at::Tensor Y, at::Tensor c, at::Tensor ratings, double reg, double lr);

at::Tensor tc_probe(at::Tensor A, at::Tensor Bm, int64_t variant);

}  // namespace gb
