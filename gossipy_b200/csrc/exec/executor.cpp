// Native executor: turns the scheduler's event list of a round into device work on per-node CUDA
// streams WITHOUT going back to Python per event.
//
// The reference walks Python objects for every message (gossipy/simul.py:389-451 -> node.py:171-204 ->
// handler.py:117-136).  For handlers whose local update is one fused kernel (1-hidden-layer MLP / logistic
// regression, SGD with or without momentum, cross-entropy) -- GossipNode, PassThroughNode, CacheNeighNode,
// SamplingBasedNode, PartitioningBasedNode, All2AllGossipNode; all four CreateModelModes -- everything an event
// needs is a handful of integers and device pointers, so the whole round is enqueued from C++:
//
//   SEND / REPLY_SEND  snapshot kernel  slot <- sender's row            on the SENDER's stream
//   DELIVER / REPLY    MERGE_UPDATE: fused merge + local-update kernel (reads the slot)
//                      UPDATE: adopt (copy) + local-update kernel;  PASS: adopt     on the RECEIVER's stream
//   partitioned models (PartitioningBasedNode + PartitionedTMH, reference node.py:566-659, handler.py:455-525):
//                      SEND also draws the partition id (keyed, like node.py does under the native engine) and
//                      carries the sender's per-partition ages; DELIVER = segment merge of that partition with
//                      age weights + local update with per-partition 1/age gradient scaling (ages by value)
//
// with two CUDA events per snapshot slot ordering the streams (slot written -> reader may start; slot
// read -> slot may be overwritten).  Disjoint node pairs therefore overlap on the device, and the host
// cost per event is a few hundred nanoseconds plus the launches.  Model ages, per-node update counters
// and the counter-based shuffle keys are the same functions as in model/handler.py, so a run is
// bit-identical to the Python executor's (tests/test_stream_executor.py).
//
// Memory is owned by Python (arena rows, one tensor of snapshot slots, the nodes' torch streams); this
// file only holds pointers.  `use_cuda = false` replaces the two launches by Python callbacks: the CPU
// tests exercise all of the bookkeeping below against the Python executor.
#include "executor.h"

#include <cuda_runtime.h>
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../kernels/kernels.h"

namespace py = pybind11;

namespace gb {

namespace {

inline uint64_t mix64(uint64_t x) {          // engine/rng.py::mix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// event kinds / message types of csrc/sched/scheduler.cpp
enum : int32_t { EV_SEND = 0, EV_DROP = 1, EV_DELIVER = 2, EV_REPLY_SEND = 3, EV_REPLY_DELIVER = 4, EV_EVAL = 5, EV_TIMEOUT = 7 };
enum : int32_t { MT_PUSH = 1, MT_PULL = 2, MT_REPLY = 3, MT_PUSH_PULL = 4 };

struct Slot {
    int64_t age = 0;
    int64_t counter = 0;                    // the sender's update counter at send time (UPDATE_MERGE keys the copy's update with it)
    int sender = -1;                        // PassThroughNode: the sender's degree rides along
    int refs = 0;                           // all-to-all mode: messages in flight + cache entries that reference this snapshot
    std::vector<cudaEvent_t> reads; int n_reads = 0;     // all-to-all mode: one event per same-rank reader of this life
    std::vector<int64_t> ages_v; int pid = 0;
    int state = 0;                          // debug mode: 0 free, 1 written (on the wire), checked on every transition
    cudaEvent_t written = nullptr, read = nullptr; bool has_reader = false;      // same-rank ordering
    float* data = nullptr;                                                        // peer-mapped when the slot lives on another rank
    uint32_t* ready = nullptr; uint32_t* done = nullptr;                          // cross-rank handshake words (owner's memory)
    uint32_t gen = 0, remote_reads = 0, acked = 0;                                // replicated bookkeeping of the handshake
};

struct Node {
    float* row = nullptr; const float* X = nullptr; const int64_t* y = nullptr;
    int n = 0; int64_t age = 0, counter = 0; cudaStream_t stream = nullptr;
    std::vector<int64_t> ages_v;            // partitioned models: one age per partition (age = their sum)
    int64_t model_msgs = 0;                 // model-carrying messages sent so far (keys the partition draw)
    float* scratch = nullptr;               // UPDATE_MERGE: private trainable copy of a received model
    uint64_t pt_draws = 0;                  // PassThroughNode: accept draws made so far (keys the next one)
    uint64_t cn_draws = 0;                  // CacheNeighNode: cache choices made so far (keys the next one)
    float* mom = nullptr; bool mom_first = false;   // fused momentum-SGD: the momentum-buffer row, still without state?
    int64_t* sample_idx = nullptr; float* sample_val = nullptr;   // SamplingTMH: the coordinate sample and its merged values
    // All2AllGossipNode: newest model per sender in first-arrival order (a Python dict's order), mixing weights ([0] = self,
    // then one per peer in get_peers() order), and the snapshot shared by the pushes of one timeout
    std::vector<std::pair<int, std::pair<int, int>>> cache;
    std::vector<double> mix_w; std::vector<int> peers;
    int64_t version = 0, snap_version = -1; int snap_rk = -1, snap_slot = -1;
    // snapshot elision: a message whose delivery provably precedes this node's next write travels as a reference to the
    // LIVE row (`alias.data == row`), no copy.  0 = none, 1 = on the wire, 2 = read: the next write waits for `alias.read`
    Slot alias; int alias_state = 0;
};

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

}  // namespace

// Several ranks (one process per GPU): every rank runs this executor over the SAME event list and keeps
// the same books (ages, counters, slot allocation, generations), but launches only the work of the nodes it
// owns.  A rank's snapshot slots live in its symmetric arena (engine/arena.py), mapped into every peer; a
// reader on another rank gets the slot's `ready` / `done` words as a PeerSync, i.e. the training kernel
// itself waits for the snapshot and acknowledges the read over NVLink (kernels/common.cuh).
class StreamExecutor {
public:
    // family 0 = mlp1 (dims IN, H, OUT), 1 = logreg (dims IN, OUT; H ignored)
    // mode: CreateModelMode value (1 UPDATE, 2 MERGE_UPDATE, 4 PASS); limited_merge >= 0 selects the age-limited
    // merge weights of LimitedMergeTMH with that threshold, -1 the uniform 0.5 / 0.5 merge
    StreamExecutor(int n_nodes, int family, int IN, int H, int OUT, int batch_size, int epochs, double lr, double wd,
                   uint64_t base_seed, bool use_cuda, int mode, int64_t limited_merge)
        : nodes_(n_nodes), owner_(n_nodes, 0), family_(family), IN_(IN), H_(H), OUT_(OUT), B_(batch_size), epochs_(epochs),
          lr_((float)lr), wd_((float)wd), seed_(base_seed), cuda_(use_cuda), mode_(mode), L_(limited_merge),
          pools_(1), free_(1) {
        if (n_nodes <= 0) throw std::invalid_argument("n_nodes must be positive");
        const char* dbg = std::getenv("GOSSIPY_EXEC_DEBUG");
        debug_ = dbg != nullptr && dbg[0] != '\0' && dbg[0] != '0';
        const char* el = std::getenv("GOSSIPY_EXEC_ELIDE");
        elide_ = !(el != nullptr && el[0] == '0');
        if (mode < 1 || mode > 4) throw std::invalid_argument("mode must be UPDATE, MERGE_UPDATE, UPDATE_MERGE or PASS");
    }
    ~StreamExecutor() {
        for (auto& pool : pools_)
            for (Slot& s : pool) {
                if (s.written) cudaEventDestroy(s.written);
                if (s.read) cudaEventDestroy(s.read);
                for (cudaEvent_t e : s.reads) cudaEventDestroy(e);
            }
        for (Node& nd : nodes_) {
            if (nd.alias.written) cudaEventDestroy(nd.alias.written);
            if (nd.alias.read) cudaEventDestroy(nd.alias.read);
        }
    }

    void set_ranks(int my_rank, int world, const std::vector<int>& owner) {
        if ((int)owner.size() != (int)nodes_.size() || my_rank < 0 || my_rank >= world)
            throw std::invalid_argument("one owner rank per node expected");
        rank_ = my_rank; world_ = world; owner_ = owner;
        pools_.resize(world); free_.resize(world);
    }
    void set_node(int i, uintptr_t row, uintptr_t X, uintptr_t y, int n, int64_t age, int64_t counter, uintptr_t stream) {
        Node& nd = nodes_.at(i);
        nd.row = reinterpret_cast<float*>(row); nd.X = reinterpret_cast<const float*>(X);
        nd.y = reinterpret_cast<const int64_t*>(y); nd.n = n; nd.age = age; nd.counter = counter;
        nd.stream = reinterpret_cast<cudaStream_t>(stream);
    }
    // PassThroughNode (node.py::_accepts): degrees of all nodes and the accept draws made so far
    void set_passthrough(const std::vector<int64_t>& degrees, const std::vector<int64_t>& draws) {
        if (degrees.size() != nodes_.size() || draws.size() != nodes_.size()) throw std::invalid_argument("one degree / counter per node expected");
        deg_.assign(degrees.begin(), degrees.end());
        for (size_t i = 0; i < nodes_.size(); ++i) nodes_[i].pt_draws = (uint64_t)draws[i];
    }
    std::vector<int64_t> pt_draws() const { std::vector<int64_t> v; for (const Node& n : nodes_) v.push_back((int64_t)n.pt_draws); return v; }
    // SamplingBasedNode + SamplingTMH (MERGE_UPDATE): the receiver draws k coordinates (keyed, with replacement), merges only
    // those, then trains (model/handler.py::SamplingTMH, node.py::SamplingBasedNode)
    void set_sampling(int64_t k, int64_t n_params) { sample_k_ = k; n_params_ = n_params; }
    void set_node_sample_buffers(int i, uintptr_t idx, uintptr_t val) {
        nodes_.at(i).sample_idx = reinterpret_cast<int64_t*>(idx); nodes_.at(i).sample_val = reinterpret_cast<float*>(val);
    }
    void set_sample_merge_callback(py::function f) { cb_sample_merge_ = std::move(f); }
    // All2AllGossipNode + WeightedTMH (node.py:406-462, handler.py::WeightedTMH): deliveries are cached per sender; on timeout the
    // cached models are merged with the mixing weights (k-way kernel) and the node trains; its pushes share one snapshot
    void set_all2all(bool on) { a2a_ = on; }
    void set_node_mixing(int i, const std::vector<int>& peers, const std::vector<double>& weights) {
        nodes_.at(i).peers = peers; nodes_.at(i).mix_w = weights;
    }
    void set_kway_callback(py::function f) { cb_kway_ = std::move(f); }
    // torch.optim.SGD with momentum inside the tensor-core kernel (MERGE_UPDATE: plain pair merge, then the momentum kernel --
    // what model/handler.py does for these handlers)
    void set_momentum(double mu, double dampening, bool nesterov) { momentum_ = (float)mu; dampening_ = (float)dampening; nesterov_ = nesterov; }
    void set_node_momentum(int i, uintptr_t buf, bool first) { nodes_.at(i).mom = reinterpret_cast<float*>(buf); nodes_.at(i).mom_first = first; }
    std::vector<int> mom_first() const { std::vector<int> v; for (const Node& n : nodes_) v.push_back(n.mom_first ? 1 : 0); return v; }
    void set_merge_pair_callback(py::function f) { cb_merge_pair_ = std::move(f); }
    // CacheNeighNode (node.py:196-226): deliveries are only stored (newest per sender); a PUSH / PUSH_PULL send first consumes
    // one cached model, chosen among the senders in the cache (sorted) by a keyed draw
    void set_cache_neigh(const std::vector<int64_t>& draws) {
        if (draws.size() != nodes_.size()) throw std::invalid_argument("one draw counter per node expected");
        cn_ = true;
        for (size_t i = 0; i < nodes_.size(); ++i) nodes_[i].cn_draws = (uint64_t)draws[i];
    }
    std::vector<int64_t> cn_draws() const { std::vector<int64_t> v; for (const Node& n : nodes_) v.push_back((int64_t)n.cn_draws); return v; }
    // (node, sender, rank, slot, age) of every cached model -- checkpointing
    std::vector<std::vector<int64_t>> caches() const {
        std::vector<std::vector<int64_t>> v;
        for (size_t i = 0; i < nodes_.size(); ++i)
            for (const auto& e : nodes_[i].cache)
                v.push_back({(int64_t)i, e.first, e.second.first, e.second.second, pools_[e.second.first][e.second.second].age});
        return v;
    }
    void import_cache(const std::vector<std::vector<int64_t>>& rows) {        // slots already filled by Python
        for (const auto& r : rows) {                                            // (node, sender, slot, age [, rank])
            const int rk = r.size() > 4 ? (int)r.at(4) : 0, s = (int)r.at(2);
            auto& fl = free_.at(rk);
            auto it = std::find(fl.begin(), fl.end(), s);
            if (it == fl.end()) throw std::invalid_argument("slot is not free");
            fl.erase(it);
            Slot& sl = pools_.at(rk).at(s);
            sl.state = 1; sl.refs = 1; sl.age = r.at(3);
            nodes_.at((size_t)r.at(0)).cache.push_back({(int)r.at(1), {rk, s}});
        }
    }
    void set_node_scratch(int i, uintptr_t scratch) { nodes_.at(i).scratch = reinterpret_cast<float*>(scratch); }
    void set_update_merge_callback(py::function f) { cb_update_merge_ = std::move(f); }
    void set_node_data(int i, uintptr_t X, uintptr_t y, int n) {       // streamed inputs: the buffers alternate per round
        Node& nd = nodes_.at(i);
        nd.X = reinterpret_cast<const float*>(X); nd.y = reinterpret_cast<const int64_t*>(y); nd.n = n;
    }
    // one rank: the snapshot slots are one [cap, stride] fp32 tensor owned by Python; growing keeps the contents (Python copies)
    void set_slots(uintptr_t base, int cap, int64_t stride_floats, int64_t row_floats) {
        auto& pool = pools_.at(0);
        if (cap < (int)pool.size()) throw std::invalid_argument("the slot pool cannot shrink");
        row_floats_ = row_floats;
        for (int s = (int)pool.size(); s < cap; ++s) { pool.emplace_back(); free_[0].push_back(s); }
        for (int s = 0; s < cap; ++s) pool[s].data = reinterpret_cast<float*>(base) + (size_t)s * stride_floats;
    }
    // several ranks: slot k of `rank`'s pool = a row of that rank's symmetric arena + its two flag words
    // (gen / remote_reads / acked continue the row's history: arena rows are recycled)
    int add_slot(int rank, uintptr_t data, uintptr_t ready, uintptr_t done, int64_t row_floats, uint32_t gen,
                 uint32_t remote_reads, uint32_t acked) {
        auto& pool = pools_.at(rank);
        Slot sl;
        sl.data = reinterpret_cast<float*>(data); sl.ready = reinterpret_cast<uint32_t*>(ready);
        sl.done = reinterpret_cast<uint32_t*>(done);
        sl.gen = gen; sl.remote_reads = remote_reads; sl.acked = acked;
        pool.push_back(sl);
        free_[rank].push_back((int)pool.size() - 1);
        row_floats_ = row_floats;
        return (int)pool.size() - 1;
    }
    void set_callbacks(py::function snapshot, py::function train, py::function adopt) {
        cb_snapshot_ = snapshot; cb_train_ = train; cb_adopt_ = adopt;
    }
    // partitioned models: partition table of every parameter (device, int64) and the strided blocks of every
    // partition (device int64 [S, 4] each, model/sampling.py::segments); CPU mode: callbacks instead
    void set_partition(int n_parts, uintptr_t part_id, const std::vector<uintptr_t>& seg_ptrs, const std::vector<int>& seg_counts) {
        if (n_parts < 1 || n_parts > kMaxPartsByValue) throw std::invalid_argument("1..16 partitions supported by the native executor");
        if ((int)seg_ptrs.size() != n_parts || (int)seg_counts.size() != n_parts) throw std::invalid_argument("one segment table per partition");
        if (mode_ != 2 && mode_ != 1) throw std::invalid_argument("partitioned models run MERGE_UPDATE or UPDATE in the native executor");
        n_parts_ = n_parts; part_id_ = reinterpret_cast<const int64_t*>(part_id);
        seg_ptrs_ = seg_ptrs; seg_counts_ = seg_counts;
        for (Node& nd : nodes_) nd.ages_v.assign(n_parts, 0);
    }
    void set_partition_callbacks(py::function merge_part, py::function train_part) { cb_merge_part_ = merge_part; cb_train_part_ = train_part; }
    void set_partition_update_callback(py::function f) { cb_update_part_ = std::move(f); }
    void set_sample_update_callback(py::function f) { cb_sample_update_ = std::move(f); }
    void set_node_ages(int i, const std::vector<int64_t>& ages, int64_t model_msgs) {
        Node& nd = nodes_.at(i);
        if ((int)ages.size() != n_parts_) throw std::invalid_argument("one age per partition");
        nd.ages_v = ages; nd.model_msgs = model_msgs;
        nd.age = 0; for (int64_t a : ages) nd.age += a;
    }
    std::vector<std::vector<int64_t>> ages_v() const { std::vector<std::vector<int64_t>> v; for (const Node& n : nodes_) v.push_back(n.ages_v); return v; }
    std::vector<int64_t> model_msgs() const { std::vector<int64_t> v; for (const Node& n : nodes_) v.push_back(n.model_msgs); return v; }
    int free_slots() const { return (int)free_[world_ > 1 ? rank_ : 0].size(); }

    // Executes one round's events [n, 6] = (kind, tick, a, b, slot, aux) from index `start`.  Returns the nodes to
    // evaluate.  When a pool runs out of slots it stops at that event (`resume_at` >= 0): the caller grows the pool
    // (one rank: set_slots) and calls again with `start = resume_at`.
    std::vector<int> run(py::array_t<int32_t, py::array::c_style | py::array::forcecast> events, int64_t start) {
        if (events.ndim() != 2 || events.shape(1) != 6) throw std::invalid_argument("events must be an [n, 6] array");
        if (!cuda_ && (!cb_snapshot_ || !cb_train_ || !cb_adopt_))
            throw std::runtime_error("CPU mode needs set_callbacks(snapshot, train, adopt)");
        if (row_floats_ <= 0) throw std::runtime_error("no snapshot slots: call set_slots / add_slot first");
        const auto ev = events.unchecked<2>();
        std::vector<int> evals;
        resume_at_ = -1;
        for (int64_t i = start; i < ev.shape(0); ++i) {
            const int32_t kind = ev(i, 0), a = ev(i, 2), b = ev(i, 3), id = ev(i, 4), aux = ev(i, 5);
            switch (kind) {
                case EV_SEND:
                    if (a2a_) { if (!snapshot_shared(a, id)) { resume_at_ = i; return evals; } break; }
                    if (cn_ && aux != MT_PULL) {
                        if (free_[world_ > 1 ? owner_[a] : 0].empty()) { resume_at_ = i; return evals; }   // (before the cache is touched)
                        consume_cached(a);
                    }
                    if (aux != MT_PULL) { if (!snapshot(a, id, can_alias(ev, i, a, b, id, false))) { resume_at_ = i; return evals; } }
                    break;
                case EV_TIMEOUT:
                    if (a2a_) on_timeout(a);
                    break;
                case EV_REPLY_SEND:                       // b answers with its (just updated) model; reply id in aux
                    if (!snapshot(b, aux, can_alias(ev, i, b, a, aux, true))) { resume_at_ = i; return evals; }
                    break;
                case EV_DROP: {
                    auto it = inflight_.find(id);
                    if (it != inflight_.end()) {
                        if (a2a_) {
                            release_ref(it->second.first, it->second.second);
                        } else if (it->second.first < 0) {       // an elided snapshot that was never read
                            Node& snd = nodes_[-1 - it->second.first];
                            snd.alias_state = 0; snd.alias.state = 0;
                        } else {
                            if (debug_) pools_[it->second.first][it->second.second].state = 0;
                            free_[it->second.first].push_back(it->second.second);
                        }
                        inflight_.erase(it);
                    }
                    break;
                }
                case EV_DELIVER:
                    if (a2a_) { if (aux == MT_PUSH) store(b, a, id); break; }
                    if (cn_) { if (aux == MT_PUSH || aux == MT_PUSH_PULL) store(b, a, id); break; }
                    if (aux == MT_PUSH || aux == MT_PUSH_PULL) consume(b, id);
                    break;
                case EV_REPLY_DELIVER:
                    if (cn_) store(a, b, id); else consume(a, id);
                    break;
                case EV_EVAL:
                    evals.push_back(a);
                    break;
                default:
                    break;
            }
        }
        return evals;
    }
    int64_t resume_at() const { return resume_at_; }
    int64_t elided() const { return elided_; }

    std::vector<int64_t> ages() const { std::vector<int64_t> v; for (const Node& n : nodes_) v.push_back(n.age); return v; }
    std::vector<int64_t> counters() const { std::vector<int64_t> v; for (const Node& n : nodes_) v.push_back(n.counter); return v; }
    int64_t launches() const { return launches_; }
    bool debug() const { return debug_; }
    // messages on the wire: (message id, rank, slot, age [, partition id, age of every partition]) -- checkpointing
    std::vector<std::vector<int64_t>> inflight() const {
        std::vector<std::vector<int64_t>> v;
        for (const auto& kv : inflight_) {
            if (kv.second.first < 0) throw std::logic_error("an elided snapshot outlived its round");
            const Slot& sl = pools_[kv.second.first][kv.second.second];
            std::vector<int64_t> r{kv.first, kv.second.first, kv.second.second, sl.age};
            if (n_parts_ > 0) {
                r.push_back(sl.pid); r.insert(r.end(), sl.ages_v.begin(), sl.ages_v.end());
                if (mode_ == 1) r.push_back(sl.counter);    // partitioned UPDATE keys the copy's update with it
            }
            else if (mode_ == 3 || (sample_k_ > 0 && mode_ == 1)) r.push_back(sl.counter);
            if (!deg_.empty()) r.push_back(sl.sender);
            v.push_back(r);
        }
        std::sort(v.begin(), v.end());
        return v;
    }
    void import_inflight(const std::vector<std::vector<int64_t>>& rows) {     // slots already filled by Python
        for (const auto& r : rows) {
            const int rk = (int)r.at(1), s = (int)r.at(2);
            auto& fl = free_.at(rk);
            auto it = std::find(fl.begin(), fl.end(), s);
            if (it == fl.end()) throw std::invalid_argument("slot is not free");
            fl.erase(it);
            Slot& sl = pools_.at(rk).at(s);
            sl.state = 1;                                   // (debug mode) holds a snapshot again
            sl.refs = 1;
            sl.age = r.at(3);
            if (n_parts_ > 0) {
                sl.pid = (int)r.at(4); sl.ages_v.assign(r.begin() + 5, r.begin() + 5 + n_parts_);
                if (mode_ == 1) sl.counter = r.at(5 + n_parts_);
            }
            else if (mode_ == 3 || (sample_k_ > 0 && mode_ == 1)) sl.counter = r.at(4);
            if (!deg_.empty()) sl.sender = (int)r.back();
            inflight_[(int32_t)r.at(0)] = {rk, s};
        }
    }

private:
    bool mine(int node) const { return world_ == 1 || owner_[node] == rank_; }
    int steps_of(const Node& nd) const {
        const int bs = B_ == 0 ? nd.n : std::min(B_, nd.n);
        return epochs_ > 0 ? epochs_ * ((nd.n + bs - 1) / bs) : 1;
    }
    uint64_t key_of(int node, int64_t counter, int64_t age) const {     // model/handler.py::_next_key -> engine/rng.py::derive
        uint64_t h = mix64(seed_);
        const uint64_t parts[4] = {0x5EEDull, (uint64_t)node, (uint64_t)counter, (uint64_t)age};
        for (uint64_t p : parts) h = mix64(h ^ p);
        return h & ((1ull << 63) - 1);
    }
    uint64_t key_of(int node, const Node& nd) const { return key_of(node, nd.counter, nd.age); }

    void fill_meta(Slot& sl, int node, Node& nd) {         // what travels with the model: age(s), counter, partition id
        sl.age = nd.age;
        sl.counter = nd.counter;
        sl.sender = node;
        if (n_parts_ > 0) {                               // node.py::PartitioningBasedNode._payload_extras (keyed form)
            sl.ages_v = nd.ages_v;
            uint64_t h = mix64(seed_);
            const uint64_t parts[3] = {0x9A57ull, (uint64_t)node, (uint64_t)nd.model_msgs};
            for (uint64_t p : parts) h = mix64(h ^ p);
            sl.pid = (int)((h & ((1ull << 63) - 1)) % (uint64_t)n_parts_);
            nd.model_msgs += 1;
        }
    }

    // Snapshot elision ("zero-copy live read when provably safe"): the message sent at event i by `sender` to `dst` may
    // travel as a reference to the sender's live row iff, in this round's event list, (1) it is delivered (or dropped)
    // before any event writes the sender's row, and (2) the sender's next write is the delivery of the REPLY to this
    // very message -- that write waits for the reader's kernel anyway (the reply is the reader's updated model), so the
    // live read adds no dependency.  Any other "next write" would have to wait for the reader's whole training kernel;
    // even "no further write in this round" is not free, because rounds pipeline on the device: the first write of the
    // next round would wait for a kernel at the tail of this round's chain (measured: eliding those replies as well
    // cost 7 % on the headline benchmark, profiles/r2d).  So: the request leg of a PUSH_PULL exchange, and messages
    // that are dropped.  Same rank only (a peer GPU reads through the published-snapshot protocol).
    template <class Ev>
    bool can_alias(const Ev& ev, int64_t i, int sender, int dst, int32_t msg_id, bool is_reply) const {
        if (!elide_ || a2a_ || cn_ || nodes_[sender].alias_state != 0) return false;
        if (world_ > 1 && owner_[sender] != owner_[dst]) return false;
        const int64_t n = ev.shape(0), window = std::min<int64_t>(n, i + 1 + 512);
        auto writes_sender = [&](int64_t j) {
            const int32_t kind = ev(j, 0);
            if (kind == EV_DELIVER) return ev(j, 3) == sender && (ev(j, 5) == MT_PUSH || ev(j, 5) == MT_PUSH_PULL);
            if (kind == EV_REPLY_DELIVER) return ev(j, 2) == sender;
            return false;
        };
        int64_t j = i + 1;
        bool found = false;
        for (; j < window; ++j) {
            const int32_t kind = ev(j, 0);
            if (kind == EV_DROP && ev(j, 4) == msg_id) return true;
            if ((!is_reply && kind == EV_DELIVER && ev(j, 4) == msg_id) || (is_reply && kind == EV_REPLY_DELIVER && ev(j, 4) == msg_id)) { found = true; break; }
            if (writes_sender(j)) return false;
        }
        if (!found) return false;
        if (is_reply) return false;
        int32_t reply_id = -1;                              // the reply this delivery triggers (PUSH_PULL): slot = request id
        for (int64_t q = j + 1; q < n && q - j <= 512; ++q) {
            const int32_t kind = ev(q, 0);
            if (kind == EV_REPLY_SEND && ev(q, 4) == msg_id) reply_id = ev(q, 5);
            if (writes_sender(q)) return kind == EV_REPLY_DELIVER && reply_id >= 0 && ev(q, 4) == reply_id;
        }
        return false;
    }

    bool snapshot(int node, int32_t msg_id, bool alias = false) {
        Node& nd = nodes_.at(node);
        if (alias) {
            Slot& sl = nd.alias;
            if (debug_ && (sl.state != 0 || nd.alias_state != 0)) throw std::logic_error("executor debug: second elided snapshot of a live row");
            sl.state = 1; nd.alias_state = 1;
            fill_meta(sl, node, nd);
            sl.data = nd.row;
            if (mine(node) && cuda_) {                      // "written" = the sender's stream has reached this point
                if (!sl.written) cuda_check(cudaEventCreateWithFlags(&sl.written, cudaEventDisableTiming), "event");
                cuda_check(cudaEventRecord(sl.written, nd.stream), "record live row ready");
            }
            inflight_[msg_id] = {-1 - node, 0};
            ++elided_;
            return true;
        }
        const int rk = world_ > 1 ? owner_[node] : 0;
        auto& fl = free_[rk];
        if (fl.empty()) return false;
        // oldest free slot first: its last reader is a whole training kernel (~1 ms) on another node's
        // stream, and the WAR wait below would serialise unrelated nodes if a just-freed slot were reused
        // (LIFO reuse measured 74 instead of 190 rounds/s on the headline benchmark)
        const int s = fl.front(); fl.pop_front();
        Slot& sl = pools_[rk][s];
        if (debug_) {       // race-debug mode (GOSSIPY_EXEC_DEBUG=1): every slot has ONE writer and ONE reader per life
            if (sl.state != 0) throw std::logic_error("executor debug: snapshot into a slot that is still on the wire");
            for (const auto& kv : inflight_)
                if (kv.second.first == rk && kv.second.second == s) throw std::logic_error("executor debug: free list and in-flight table share a slot");
            if (inflight_.count(msg_id)) throw std::logic_error("executor debug: message id sent twice");
            if (world_ > 1 && sl.acked > sl.remote_reads) throw std::logic_error("executor debug: more acknowledgements than remote reads");
            sl.state = 1;
        }
        fill_meta(sl, node, nd);
        sl.gen += 1;                                      // replicated: every rank knows which generation a reader expects
        if (mine(node)) {
            if (cuda_) {
                // the slot's previous life: its reader (WAR) and -- for a dropped message nobody read -- its writer (WAW);
                // readers on other ranks acknowledge through the slot's `done` counter
                if (sl.has_reader) cuda_check(cudaStreamWaitEvent(nd.stream, sl.read, 0), "wait for the slot's last reader");
                for (int q = 0; q < sl.n_reads; ++q) cuda_check(cudaStreamWaitEvent(nd.stream, sl.reads[q], 0), "wait for the slot's readers");
                sl.n_reads = 0;
                if (sl.written) cuda_check(cudaStreamWaitEvent(nd.stream, sl.written, 0), "wait for the slot's last writer");
                if (sl.remote_reads != sl.acked) launch_flag_wait(sl.done, sl.remote_reads, nd.stream);
                launch_merge_pair(sl.data, nd.row, 0.f, 1.f, 0, row_floats_, PeerSync{nullptr, 0, nullptr, nullptr}, nd.stream);
                if (!sl.written) cuda_check(cudaEventCreateWithFlags(&sl.written, cudaEventDisableTiming), "event");
                cuda_check(cudaEventRecord(sl.written, nd.stream), "record slot written");
                if (world_ > 1) launch_flag_signal(sl.ready, sl.gen, nd.stream);
                cuda_check(cudaGetLastError(), "snapshot launch");
            } else {
                cb_snapshot_(node, rk, s, (int64_t)sl.gen, (int64_t)sl.remote_reads);
            }
            ++launches_;
        }
        sl.acked = sl.remote_reads;
        inflight_[msg_id] = {rk, s};
        return true;
    }

    // ---- All2AllGossipNode --------------------------------------------------------------------------------------------
    void release_ref(int rk, int s) {
        Slot& sl = pools_[rk][s];
        if (--sl.refs <= 0) { sl.refs = 0; sl.state = 0; free_[rk].push_back(s); }
    }
    // the pushes of one timeout carry the same model: one snapshot, one reference per message (model/handler.py::caching)
    bool snapshot_shared(int node, int32_t msg_id) {
        Node& nd = nodes_.at(node);
        if (nd.snap_slot >= 0 && nd.snap_version == nd.version && pools_[nd.snap_rk][nd.snap_slot].refs > 0) {
            pools_[nd.snap_rk][nd.snap_slot].refs += 1;
            if (debug_ && inflight_.count(msg_id)) throw std::logic_error("executor debug: message id sent twice");
            inflight_[msg_id] = {nd.snap_rk, nd.snap_slot};
            return true;
        }
        if (!snapshot(node, msg_id, false)) return false;
        const auto where = inflight_.at(msg_id);
        pools_[where.first][where.second].refs = 1;
        nd.snap_rk = where.first; nd.snap_slot = where.second; nd.snap_version = nd.version;
        return true;
    }
    void store(int node, int sender, int32_t msg_id) {          // node.py::All2AllGossipNode.receive: newest model per sender
        auto it = inflight_.find(msg_id);
        if (it == inflight_.end()) throw std::runtime_error("delivery of an unknown message");
        const std::pair<int, int> where = it->second;
        inflight_.erase(it);
        Node& nd = nodes_.at(node);
        for (auto& e : nd.cache)
            if (e.first == sender) { release_ref(e.second.first, e.second.second); e.second = where; return; }
        nd.cache.push_back({sender, where});
    }
    void consume_cached(int node) {                               // node.py::CacheNeighNode.send: one cached model, keyed choice
        Node& nd = nodes_.at(node);
        if (nd.cache.empty()) return;
        std::vector<int> keys;
        for (const auto& e : nd.cache) keys.push_back(e.first);
        std::sort(keys.begin(), keys.end());
        uint64_t h = mix64(seed_);
        const uint64_t parts[3] = {0x9A59ull, (uint64_t)node, nd.cn_draws++};
        for (uint64_t p : parts) h = mix64(h ^ p);
        const int pick = keys[(size_t)((h & ((1ull << 63) - 1)) % (uint64_t)keys.size())];
        for (size_t q = 0; q < nd.cache.size(); ++q)
            if (nd.cache[q].first == pick) {
                const std::pair<int, int> where = nd.cache[q].second;
                nd.cache.erase(nd.cache.begin() + (long)q);
                pools_[where.first][where.second].refs = 0;
                consume_at(node, where.first, where.second);
                return;
            }
    }
    void on_timeout(int node) {                                   // node.py::All2AllGossipNode.on_timeout + WeightedTMH MERGE_UPDATE
        Node& nd = nodes_.at(node);
        if (nd.cache.empty()) return;
        const int k = (int)nd.cache.size();
        std::vector<double> use(1, nd.mix_w.empty() ? 0.0 : nd.mix_w[0]);
        for (const auto& e : nd.cache) {
            double w = 0.0;
            for (size_t q = 0; q < nd.peers.size(); ++q)
                if (nd.peers[q] == e.first) { if (q + 1 < nd.mix_w.size()) w = nd.mix_w[q + 1]; break; }
            use.push_back(w);
        }
        if (k < (int)nd.peers.size()) {                           // fewer models than neighbours: renormalise (sequential sum,
            double sum = 0.0;                                     // like the Python side)
            for (double u : use) sum += u;
            if (sum > 0.0) for (double& u : use) u /= sum;
        }
        const bool exec = mine(node);
        int64_t age = nd.age;
        std::vector<const float*> srcs; std::vector<PeerSync> syncs; bool any_remote = false;
        std::vector<std::vector<int64_t>> cpu_srcs;
        for (const auto& e : nd.cache) {
            Slot& sl = pools_[e.second.first][e.second.second];
            age = std::max(age, sl.age);
            const bool remote = world_ > 1 && e.second.first != owner_[node];
            if (remote) sl.remote_reads += 1;                     // replicated: the owner waits for this many acknowledgements
            if (exec && cuda_) {
                srcs.push_back(sl.data);
                if (remote) { syncs.push_back(PeerSync{sl.ready, sl.gen, sl.done, device_fault_word()}); any_remote = true; }
                else {
                    syncs.push_back(PeerSync{nullptr, 0, nullptr, nullptr});
                    if (sl.written) cuda_check(cudaStreamWaitEvent(nd.stream, sl.written, 0), "wait for a cached snapshot");
                }
            } else if (exec) {
                cpu_srcs.push_back({e.second.first, e.second.second, (int64_t)sl.gen});
            }
        }
        if (exec) {
            if (cuda_) {
                std::vector<float> wf(use.begin(), use.end());
                launch_merge_kway(nd.row, srcs.data(), wf.data(), k, row_floats_, any_remote ? syncs.data() : nullptr, nd.stream);
                for (const auto& e : nd.cache) {                  // same-rank readers: the slot's next writer waits for this kernel
                    if (world_ > 1 && e.second.first != owner_[node]) continue;
                    Slot& sl = pools_[e.second.first][e.second.second];
                    if ((int)sl.reads.size() <= sl.n_reads) {
                        cudaEvent_t ev; cuda_check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event");
                        sl.reads.push_back(ev);
                    }
                    cuda_check(cudaEventRecord(sl.reads[sl.n_reads++], nd.stream), "record slot read");
                }
            } else {
                cb_kway_(node, cpu_srcs, use);
            }
            ++launches_;
        }
        nd.age = age;
        nd.counter += 1;
        const uint64_t key = key_of(node, nd);
        if (exec) {
            if (cuda_) train(nd, nullptr, 1.f, 0.f, key, PeerSync{nullptr, 0, nullptr, nullptr});
            else cb_train_(node, 0, -1, (int64_t)key, 1.f, 0.f, (int64_t)0);
            ++launches_;
            if (cuda_) cuda_check(cudaGetLastError(), "timeout launch");
        }
        nd.age += steps_of(nd);
        nd.version += 1;
        for (const auto& e : nd.cache) release_ref(e.second.first, e.second.second);
        nd.cache.clear();
    }

    void merge_weights(int64_t a, int64_t b, float& ws, float& wp) const {   // model/handler.py: _fused_merge_weights
        if (L_ < 0) { ws = 0.5f; wp = 0.5f; return; }
        if (a > b + L_) { ws = 1.f; wp = 0.f; return; }
        if (b > a + L_) { ws = 0.f; wp = 1.f; return; }
        const int64_t tot = a + b;
        if (tot == 0) { ws = 0.5f; wp = 0.5f; return; }
        ws = (float)((double)a / (double)tot); wp = (float)((double)b / (double)tot);
    }

    // stage_mode (kernels.h): 1 = issue only the operand loader (returns whether anything was staged), 2 = the
    // training kernel on operands staged by an earlier call, 0 = both
    bool train(Node& nd, const float* peer, float ws, float wp, uint64_t key, PeerSync sync,
               const std::vector<int64_t>* part_ages = nullptr, int stage_mode = 0) {
        bool ok;
        const char* why = "";
        auto scaled = [&](auto& p) {
            if (part_ages == nullptr) return;
            p.part_id = part_id_; p.ages = nullptr; p.use_ages_val = true; p.n_parts = n_parts_;
            for (int i = 0; i < n_parts_; ++i) p.ages_val[i] = (*part_ages)[i];
        };
        if (family_ == 0) {
            TrainParams p{};
            scaled(p);
            p.row = nd.row; p.X = nd.X; p.y = nd.y; p.n = nd.n; p.IN = IN_; p.H = H_; p.OUT = OUT_;
            p.B = B_ == 0 ? nd.n : std::min(B_, nd.n); p.epochs = epochs_; p.lr = lr_; p.wd = wd_; p.key = key;
            if (peer) { p.peer = peer; p.w_self = ws; p.w_peer = wp; p.sync = sync; }
            if (momentum_ != 0.f) {
                if (nd.mom == nullptr) throw std::runtime_error("fused momentum-SGD needs a momentum row per node (set_node_momentum)");
                p.momentum = momentum_; p.dampening = dampening_; p.nesterov = nesterov_; p.mom = nd.mom; p.mom_first = nd.mom_first;
            }
            p.stage_mode = stage_mode;
            ok = launch_mlp1_train(p, kTrainAuto, nd.stream, &why);
            if (stage_mode == 1) return ok;
            nd.mom_first = false;
        } else {
            if (stage_mode == 1) return false;
            LogregParams p{};
            scaled(p);
            p.row = nd.row; p.X = nd.X; p.y = nd.y; p.n = nd.n; p.IN = IN_; p.OUT = OUT_;
            p.B = B_ == 0 ? nd.n : std::min(B_, nd.n); p.epochs = epochs_; p.lr = lr_; p.wd = wd_; p.key = key;
            if (peer) { p.peer = peer; p.w_self = ws; p.w_peer = wp; p.sync = sync; }
            ok = launch_logreg_train(p, nd.stream);
        }
        if (!ok) throw std::runtime_error(std::string("training kernel rejected the shape: ") + why);
        return true;
    }

    void consume(int node, int32_t msg_id) {
        auto it = inflight_.find(msg_id);
        if (it == inflight_.end()) throw std::runtime_error("delivery of an unknown message");
        const int rk = it->second.first, s = it->second.second;
        inflight_.erase(it);
        consume_at(node, rk, s);
    }
    // node `node` consumes the snapshot in slot (rk, s) according to the CreateModelMode (rk < 0: an elided snapshot)
    void consume_at(int node, int rk, int s) {
        Node& nd = nodes_.at(node);
        const bool aliased = rk < 0;                        // elided snapshot: the "slot" is the sender's live row
        Slot& sl = aliased ? nodes_[-1 - rk].alias : pools_[rk][s];
        if (nd.alias_state == 1) throw std::logic_error("a node is written while its live row is on the wire (elision look-ahead violated)");
        // my live row was read by a peer's kernel: that kernel must finish before this stream WRITES the row (the operand
        // loader below does not touch the row and is issued first)
        bool wait_reader = false;
        if (nd.alias_state == 2) { wait_reader = mine(node) && cuda_ && nd.alias.has_reader; nd.alias_state = 0; }
        auto reader_done = [&]() {
            if (wait_reader) cuda_check(cudaStreamWaitEvent(nd.stream, nd.alias.read, 0), "wait for the reader of the live row");
            wait_reader = false;
        };
        if (debug_) {
            if (sl.state != 1) throw std::logic_error("executor debug: delivery of a slot that holds no snapshot");
            if (cuda_ && mine(node) && !(world_ > 1 && rk != owner_[node]) && sl.written == nullptr && sl.gen == 0)
                throw std::logic_error("executor debug: local reader without a writer event");
            sl.state = 0;
        }
        const bool exec = mine(node);
        const bool remote = !aliased && world_ > 1 && rk != owner_[node];      // the snapshot lives on another rank than the reader
        if (remote) sl.remote_reads += 1;                           // replicated: the owner will wait for this many acks
        PeerSync sync{nullptr, 0, nullptr, nullptr};
        int mode = mode_;
        if (!deg_.empty()) {                                // PassThroughNode: merge with probability min(1, deg_sender / deg_self),
            const uint64_t k = nd.pt_draws++;               // else adopt untouched (PASS); keyed draw, integer comparison
            uint64_t h = mix64(seed_);
            const uint64_t parts[3] = {0x9A55ull, (uint64_t)node, k};
            for (uint64_t p : parts) h = mix64(h ^ p);
            const uint64_t u = (h & ((1ull << 63) - 1)) >> 20;
            if (!(u * (uint64_t)deg_[node] < ((uint64_t)deg_[sl.sender] << 43))) mode = 4;
        }
        // fused MERGE_UPDATE of the MLP: the operand loader of the training kernel does not depend on the incoming model,
        // so it is issued BEFORE this stream waits for the snapshot (off the critical path of a gossip chain)
        const bool hoist = exec && cuda_ && !remote && sl.written && n_parts_ == 0 && sample_k_ == 0 && mode == 2 && family_ == 0 && momentum_ == 0.f;
        if (!hoist) reader_done();
        if (exec && cuda_) {
            if (remote) sync = PeerSync{sl.ready, sl.gen, sl.done, device_fault_word()};
            else if (sl.written && !hoist)                      // (a slot restored from a checkpoint has no writer event)
                cuda_check(cudaStreamWaitEvent(nd.stream, sl.written, 0), "wait for the snapshot");
        }
        if (n_parts_ > 0 && mode == 1) {
            // partitioned UPDATE (model/handler.py::PartitionedTMH.__call__, reference handler.py:497-506): a private copy of
            // the received model is trained on the own data -- with ITS per-partition ages scaling the gradient, keyed like
            // the Python scratch copy (this node, the sender's counter + 1, the sender's total age) --, then partition `pid`
            // of the copy is merged into the own model with the age weights.  The own model is not trained.
            const int pid = sl.pid;
            const int st = steps_of(nd);
            int64_t age_tmp = 0; for (int64_t v : sl.ages_v) age_tmp += v;
            const uint64_t key_tmp = key_of(node, sl.counter + 1, age_tmp);
            const int64_t a = nd.ages_v[pid], b = sl.ages_v[pid] + st;
            float w1 = .5f, w2 = .5f;
            if (a + b > 0) { w1 = (float)((double)a / (double)(a + b)); w2 = (float)((double)b / (double)(a + b)); }
            if (exec) {
                if (cuda_) {
                    if (nd.scratch == nullptr) throw std::runtime_error("partitioned UPDATE needs a scratch row per node (set_node_scratch)");
                    const PeerSync none{nullptr, 0, nullptr, nullptr};
                    launch_merge_pair(nd.scratch, sl.data, 0.f, 1.f, 0, row_floats_, sync, nd.stream);
                    Node tmp = nd; tmp.row = nd.scratch;
                    train(tmp, nullptr, 1.f, 0.f, key_tmp, none, &sl.ages_v);
                    launch_merge_segments(nd.row, nd.scratch, reinterpret_cast<const int64_t*>(seg_ptrs_[pid]), seg_counts_[pid], w1, w2, none, nd.stream);
                } else {
                    cb_update_part_(node, rk, s, (int64_t)sl.gen, (int64_t)key_tmp, sl.ages_v, pid, w1, w2);
                }
                launches_ += 3;
            }
            nd.ages_v[pid] = std::max(a, b);
            nd.age = 0; for (int64_t v : nd.ages_v) nd.age += v;
            nd.version += 1;
        } else if (n_parts_ > 0) {                         // partitioned MERGE_UPDATE: merge one partition, then train
            const int pid = sl.pid;
            const int64_t a = nd.ages_v[pid], b = sl.ages_v[pid];
            float w1 = .5f, w2 = .5f;                      // sampling.py::mixing_weights: (0, 0) -> (1, 1)
            if (a + b > 0) { w1 = (float)((double)a / (double)(a + b)); w2 = (float)((double)b / (double)(a + b)); }
            if (exec) {
                if (cuda_) launch_merge_segments(nd.row, sl.data, reinterpret_cast<const int64_t*>(seg_ptrs_[pid]), seg_counts_[pid], w1, w2, sync, nd.stream);
                else cb_merge_part_(node, rk, s, pid, w1, w2, (int64_t)sl.gen);
                ++launches_;
            }
            nd.ages_v[pid] = std::max(a, b);
            nd.age = 0; for (int64_t v : nd.ages_v) nd.age += v;
            nd.counter += 1;
            const uint64_t key = key_of(node, nd);
            if (exec) {
                if (cuda_) train(nd, nullptr, 1.f, 0.f, key, PeerSync{nullptr, 0, nullptr, nullptr}, &nd.ages_v);
                else cb_train_part_(node, (int64_t)key, nd.ages_v);
                ++launches_;
            }
            const int st = steps_of(nd);
            for (int64_t& v : nd.ages_v) v += st;
            nd.age += (int64_t)st * n_parts_;
        } else if (sample_k_ > 0 && mode == 1) {
            // sampled UPDATE (model/handler.py::SamplingTMH.__call__, reference handler.py:440-452): the receiver draws the
            // coordinate sample (one of its update keys), trains a private copy of the received model on its data (keyed
            // like the Python scratch copy) and merges the sampled coordinates of the copy; its own age does not move
            nd.counter += 1;
            uint64_t h = mix64(seed_);
            const uint64_t parts[3] = {0x5A3Full, (uint64_t)node, key_of(node, nd) & 0xFFFFFFFFull};
            for (uint64_t p : parts) h = mix64(h ^ p);
            const uint64_t key_s = h & ((1ull << 63) - 1);
            const uint64_t key_tmp = key_of(node, sl.counter + 1, sl.age);
            if (exec) {
                if (cuda_) {
                    if (nd.sample_idx == nullptr || nd.scratch == nullptr) throw std::runtime_error("sampled UPDATE needs sample buffers and a scratch row per node");
                    const PeerSync none{nullptr, 0, nullptr, nullptr};
                    launch_keyed_randint(nd.sample_idx, sample_k_, n_params_, key_s, nd.stream);
                    launch_merge_pair(nd.scratch, sl.data, 0.f, 1.f, 0, row_floats_, sync, nd.stream);
                    Node tmp = nd; tmp.row = nd.scratch;
                    train(tmp, nullptr, 1.f, 0.f, key_tmp, none);
                    launch_merge_indexed(nd.row, nd.scratch, nd.sample_idx, sample_k_, .5f, .5f, nd.sample_val, none, nd.stream);
                } else {
                    cb_sample_update_(node, rk, s, (int64_t)sl.gen, (int64_t)key_s, (int64_t)key_tmp);
                }
                launches_ += 5;
            }
        } else if (sample_k_ > 0) {                        // sampled MERGE_UPDATE: merge k keyed coordinates, then train
            nd.counter += 1;                               // SamplingTMH.draw_sample consumes one update key ...
            uint64_t h = mix64(seed_);
            const uint64_t parts[3] = {0x5A3Full, (uint64_t)node, key_of(node, nd) & 0xFFFFFFFFull};
            for (uint64_t p : parts) h = mix64(h ^ p);
            const uint64_t key_s = h & ((1ull << 63) - 1);
            if (exec) {
                if (cuda_) {
                    if (nd.sample_idx == nullptr) throw std::runtime_error("sampling needs per-node sample buffers");
                    launch_keyed_randint(nd.sample_idx, sample_k_, n_params_, key_s, nd.stream);
                    launch_merge_indexed(nd.row, sl.data, nd.sample_idx, sample_k_, .5f, .5f, nd.sample_val, sync, nd.stream);
                } else {
                    cb_sample_merge_(node, rk, s, (int64_t)key_s, (int64_t)sl.gen);
                }
                launches_ += 3;
            }
            nd.counter += 1;                               // ... and the local update the next one
            const uint64_t key = key_of(node, nd);
            if (exec) {
                if (cuda_) train(nd, nullptr, 1.f, 0.f, key, PeerSync{nullptr, 0, nullptr, nullptr});
                else cb_train_(node, rk, -1, (int64_t)key, 1.f, 0.f, (int64_t)sl.gen);
                ++launches_;
            }
            nd.age += steps_of(nd);
        } else if (mode == 3) {
            // UPDATE_MERGE (model/handler.py::__call__, reference handler.py:129-132): update the own model, update a
            // private copy of the received one on the own data (keyed like the Python scratch copy: this node, the
            // sender's counter + 1, the sender's age), then merge the two
            const int st = steps_of(nd);
            nd.counter += 1;
            const uint64_t key_own = key_of(node, nd);
            const uint64_t key_tmp = key_of(node, sl.counter + 1, sl.age);
            const int64_t age_own = nd.age + st, age_tmp = sl.age + st;
            float ws, wp;
            merge_weights(age_own, age_tmp, ws, wp);
            if (exec) {
                if (cuda_) {
                    if (nd.scratch == nullptr) throw std::runtime_error("UPDATE_MERGE needs a scratch row per node (set_node_scratch)");
                    const PeerSync none{nullptr, 0, nullptr, nullptr};
                    train(nd, nullptr, 1.f, 0.f, key_own, none);
                    launch_merge_pair(nd.scratch, sl.data, 0.f, 1.f, 0, row_floats_, sync, nd.stream);
                    Node tmp = nd; tmp.row = nd.scratch;
                    train(tmp, nullptr, 1.f, 0.f, key_tmp, none);
                    if (wp != 0.f) launch_merge_pair(nd.row, nd.scratch, ws, wp, 0, row_floats_, none, nd.stream);
                } else {
                    cb_update_merge_(node, rk, s, (int64_t)key_own, (int64_t)key_tmp, ws, wp, (int64_t)sl.gen);
                }
                launches_ += 4;
            }
            nd.age = std::max(age_own, age_tmp);
        } else if (mode == 4) {                            // PASS: adopt the received model, age unchanged
            if (exec) {
                if (cuda_) launch_merge_pair(nd.row, sl.data, 0.f, 1.f, 0, row_floats_, sync, nd.stream);
                else cb_adopt_(node, rk, s, (int64_t)sl.gen);
                ++launches_;
            }
        } else {
            float ws = 0.f, wp = 1.f;
            const bool fused_merge = mode == 2 && momentum_ == 0.f;   // MERGE_UPDATE: the merge rides on the training kernel
            if (fused_merge) {
                merge_weights(nd.age, sl.age, ws, wp);
                nd.age = std::max(nd.age, sl.age);
            } else if (mode == 2) {                        // momentum-SGD: pair merge, then the momentum kernel (handler.py: _merge, _update)
                merge_weights(nd.age, sl.age, ws, wp);
                nd.age = std::max(nd.age, sl.age);
                if (wp == 0.f) {                           // the own model is much older: nothing is read (nor acknowledged)
                    if (remote) sl.remote_reads -= 1;
                } else if (exec) {
                    if (cuda_) launch_merge_pair(nd.row, sl.data, ws, wp, 0, row_floats_, sync, nd.stream);
                    else cb_merge_pair_(node, rk, s, ws, wp, (int64_t)sl.gen);
                    ++launches_;
                }
            } else {                                       // UPDATE: adopt (a true copy: heals a diverged model), then train
                if (exec) {
                    if (cuda_) launch_merge_pair(nd.row, sl.data, 0.f, 1.f, 0, row_floats_, sync, nd.stream);
                    else cb_adopt_(node, rk, s, (int64_t)sl.gen);
                    ++launches_;
                }
                nd.age = sl.age;
            }
            nd.counter += 1;
            const uint64_t key = key_of(node, nd);
            if (exec) {
                if (hoist) {
                    const bool staged = train(nd, sl.data, ws, wp, key, sync, nullptr, 1);
                    reader_done();
                    cuda_check(cudaStreamWaitEvent(nd.stream, sl.written, 0), "wait for the snapshot");
                    train(nd, sl.data, ws, wp, key, sync, nullptr, staged ? 2 : 0);
                } else if (cuda_) train(nd, fused_merge ? sl.data : nullptr, ws, wp, key, sync);
                else cb_train_(node, rk, fused_merge ? s : -1, (int64_t)key, ws, wp, (int64_t)sl.gen);
                ++launches_;
            }
            nd.age += steps_of(nd);
        }
        if (exec && cuda_ && !remote) {
            if (!sl.read) cuda_check(cudaEventCreateWithFlags(&sl.read, cudaEventDisableTiming), "event");
            cuda_check(cudaEventRecord(sl.read, nd.stream), "record slot read");
            sl.has_reader = true;
        }
        if (exec && cuda_) cuda_check(cudaGetLastError(), "consume launch");
        if (aliased) nodes_[-1 - rk].alias_state = 2;
        else free_[rk].push_back(s);
    }

    std::vector<Node> nodes_;
    std::vector<int> owner_;
    int family_, IN_, H_, OUT_, B_, epochs_;
    float lr_, wd_;
    uint64_t seed_;
    bool cuda_;
    int mode_; int64_t L_;
    int rank_ = 0, world_ = 1;
    int64_t row_floats_ = 0;
    std::vector<std::vector<Slot>> pools_;          // per owner rank
    std::vector<std::deque<int>> free_;             // FIFO: a slot is reused as late as possible (see snapshot())
    std::unordered_map<int32_t, std::pair<int, int>> inflight_;     // message id -> (rank, slot)
    py::function cb_snapshot_, cb_train_, cb_adopt_, cb_merge_part_, cb_train_part_, cb_update_part_, cb_update_merge_, cb_sample_merge_, cb_sample_update_, cb_kway_;
    bool a2a_ = false;                               // All2AllGossipNode mode
    bool cn_ = false;                                // CacheNeighNode mode
    float momentum_ = 0.f, dampening_ = 0.f; bool nesterov_ = false;     // fused momentum-SGD (0 = plain SGD)
    py::function cb_merge_pair_;
    int64_t sample_k_ = 0, n_params_ = 0;            // SamplingTMH: sample size (0 = whole-model merges)
    int n_parts_ = 0; const int64_t* part_id_ = nullptr;
    std::vector<int64_t> deg_;                       // PassThroughNode: node degrees (empty = plain nodes)
    std::vector<uintptr_t> seg_ptrs_; std::vector<int> seg_counts_;
    int64_t launches_ = 0, resume_at_ = -1;
    bool debug_ = false, elide_ = true;
    int64_t elided_ = 0;
};

void bind_executor(py::module_& m) {
    py::class_<StreamExecutor>(m, "StreamExecutor", "Native executor of a round's event list (csrc/exec/executor.cpp)")
        .def(py::init<int, int, int, int, int, int, int, double, double, uint64_t, bool, int, int64_t>(), py::arg("n_nodes"),
             py::arg("family"), py::arg("IN"), py::arg("H"), py::arg("OUT"), py::arg("batch_size"), py::arg("epochs"),
             py::arg("lr"), py::arg("wd"), py::arg("base_seed"), py::arg("use_cuda"), py::arg("mode") = 2,
             py::arg("limited_merge") = -1)
        .def("set_ranks", &StreamExecutor::set_ranks)
        .def("add_slot", &StreamExecutor::add_slot)
        .def("set_node", &StreamExecutor::set_node)
        .def("set_node_data", &StreamExecutor::set_node_data)
        .def("set_node_scratch", &StreamExecutor::set_node_scratch)
        .def("set_all2all", &StreamExecutor::set_all2all)
        .def("set_node_mixing", &StreamExecutor::set_node_mixing)
        .def("set_kway_callback", &StreamExecutor::set_kway_callback)
        .def("set_momentum", &StreamExecutor::set_momentum)
        .def("set_node_momentum", &StreamExecutor::set_node_momentum)
        .def("mom_first", &StreamExecutor::mom_first)
        .def("set_merge_pair_callback", &StreamExecutor::set_merge_pair_callback)
        .def("set_cache_neigh", &StreamExecutor::set_cache_neigh)
        .def("cn_draws", &StreamExecutor::cn_draws)
        .def("caches", &StreamExecutor::caches)
        .def("import_cache", &StreamExecutor::import_cache)
        .def("set_sampling", &StreamExecutor::set_sampling)
        .def("set_node_sample_buffers", &StreamExecutor::set_node_sample_buffers)
        .def("set_sample_merge_callback", &StreamExecutor::set_sample_merge_callback)
        .def("set_passthrough", &StreamExecutor::set_passthrough)
        .def("pt_draws", &StreamExecutor::pt_draws)
        .def("set_update_merge_callback", &StreamExecutor::set_update_merge_callback)
        .def("set_slots", &StreamExecutor::set_slots)
        .def("set_callbacks", &StreamExecutor::set_callbacks)
        .def("set_partition", &StreamExecutor::set_partition)
        .def("set_partition_callbacks", &StreamExecutor::set_partition_callbacks)
        .def("set_partition_update_callback", &StreamExecutor::set_partition_update_callback)
        .def("set_sample_update_callback", &StreamExecutor::set_sample_update_callback)
        .def("set_node_ages", &StreamExecutor::set_node_ages)
        .def("ages_v", &StreamExecutor::ages_v)
        .def("model_msgs", &StreamExecutor::model_msgs)
        .def("run", &StreamExecutor::run, py::arg("events"), py::arg("start") = 0)
        .def_property_readonly("resume_at", &StreamExecutor::resume_at)
        .def_property_readonly("elided", &StreamExecutor::elided)
        .def_property_readonly("free_slots", &StreamExecutor::free_slots)
        .def_property_readonly("launches", &StreamExecutor::launches)
        .def_property_readonly("debug", &StreamExecutor::debug)
        .def("ages", &StreamExecutor::ages)
        .def("counters", &StreamExecutor::counters)
        .def("inflight", &StreamExecutor::inflight)
        .def("import_inflight", &StreamExecutor::import_inflight);
}

}  // namespace gb
