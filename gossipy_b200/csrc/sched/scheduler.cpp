// Native discrete-time gossip scheduler: the control plane of the simulators' round loop in C++.
//
// The reference spends every timestep of every round in Python: it scans all N nodes for
// timeouts, draws N availability samples, walks two dict-of-list message queues
// (gossipy/simul.py:389-451).  Here one call simulates a whole round (delta timesteps) natively
// and returns the EVENT LIST of that round -- sends (snapshots), deliveries, replies, drops,
// evaluations -- in exactly the order the reference's four phases would have produced them:
//
//   phase A  for nodes in the round's shuffled order: timeout? -> (token account) -> pick a peer,
//            SEND, drop test `>=`, delay, enqueue                      (simul.py:393-407, 602-615)
//   phase B  availability draw for every node; deliveries due at t in queue order, incl. messages
//            enqueued at t with delay 0 and reactive sends             (simul.py:409-421, 617-648)
//            a delivered PULL / PUSH_PULL produces a reply: drop test `>`, delay, enqueue
//   phase C  replies due at t                                          (simul.py:423-430)
//   phase D  at the end of the round: evaluation sample                (simul.py:432-450)
//
// The executor (Python for arbitrary handlers, see simul.py::_run_native) turns events into device
// work on per-node CUDA streams; the scheduler itself never touches the GPU, so every rank of a
// multi-GPU run computes the identical schedule from the seed alone (no control traffic).
// Token accounts (gossipy/flow_control.py) are evaluated natively with a constant utility.
// Randomness: counter-based splitmix64 streams keyed by (seed, purpose), independent of Python RNGs.
#include "scheduler.h"

#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <map>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

namespace gb {

namespace {

inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct Stream {   // one independent random stream
    uint64_t key, ctr;
    Stream() : key(0), ctr(0) {}
    Stream(uint64_t seed, uint64_t purpose) : key(mix64(mix64(seed) ^ purpose)), ctr(0) {}
    uint64_t next() { return mix64(key ^ (ctr++ * 0xD1342543DE82EF95ull)); }
    uint64_t at(uint64_t c) const { return mix64(key ^ (c * 0xD1342543DE82EF95ull)); }      // the draw `next()` makes at ctr == c
    double uniform_at(uint64_t c) const { return (double)(at(c) >> 11) * (1.0 / 9007199254740992.0); }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0, 1)
    uint64_t below(uint64_t n) { return n <= 1 ? 0 : next() % n; }
};

enum EventKind : int32_t {
    EV_SEND = 0,           // a = sender, b = receiver, slot = message id, aux = message type
    EV_DROP = 1,           // message `slot` is lost (dropped on the wire or receiver offline)
    EV_DELIVER = 2,        // b receives message `slot` (sent by a)
    EV_REPLY_SEND = 3,     // b answers a's PULL / PUSH_PULL: reply message id in aux (same event), slot = request
    EV_REPLY_DELIVER = 4,  // a receives reply `slot` from b
    EV_EVAL = 5,           // evaluate node a (end of round)
    EV_TOKEN = 6,          // node a banked a token instead of sending (informational)
    EV_TIMEOUT = 7,        // node a timed out (emitted before its sends when `emit_timeouts` is on)
};
enum Protocol : int32_t { PUSH = 1, PULL = 2, PUSH_PULL = 3 };
enum MsgType : int32_t { MT_PUSH = 1, MT_PULL = 2, MT_REPLY = 3, MT_PUSH_PULL = 4 };
enum DelayKind : int32_t { DELAY_CONST = 0, DELAY_UNIFORM = 1, DELAY_LINEAR = 2 };
enum TokenKind : int32_t { TK_NONE = 0, TK_PROACTIVE = 1, TK_REACTIVE = 2, TK_SIMPLE = 3, TK_GENERALIZED = 4,
                           TK_RANDOMIZED = 5 };

struct Msg { int32_t id, sender, receiver, type; int64_t size; };

struct TokenAccount {   // gossipy/flow_control.py
    int kind = TK_NONE; int64_t C = 1, A = 1, k = 1; int64_t n = 0;
    void add(int64_t v) { n += v; }
    void sub(int64_t v) { n = std::max<int64_t>(0, n - v); }
    double proactive() const {
        switch (kind) {
            case TK_PROACTIVE: return 1.0;
            case TK_REACTIVE: return 0.0;
            case TK_SIMPLE: case TK_GENERALIZED: return n >= C ? 1.0 : 0.0;
            case TK_RANDOMIZED:
                if (n < A - 1) return 0.0;
                if (n <= C) return (double)(n - A + 1) / (double)(C - A + 1);
                return 1.0;
            default: return 1.0;
        }
    }
    int64_t reactive(int64_t utility, Stream& rng) const {
        switch (kind) {
            case TK_PROACTIVE: return 0;
            case TK_REACTIVE: return utility * k;
            case TK_SIMPLE: return n > 0 ? 1 : 0;
            case TK_GENERALIZED: return utility > 0 ? (A - 1 + n) / A : (A - 1 + n) / (2 * A);
            case TK_RANDOMIZED: {
                if (utility <= 0) return 0;
                const double r = (double)n / (double)A;
                const int64_t ip = (int64_t)r;
                return ip + (rng.uniform() < (r - (double)ip) ? 1 : 0);
            }
            default: return 0;
        }
    }
};

}  // namespace

class GossipScheduler {
public:
    GossipScheduler(int n_nodes, int delta, int protocol, double drop_prob, double online_prob,
                    double sampling_eval, uint64_t seed)
        : n_(n_nodes), delta_(delta), protocol_(protocol), drop_(drop_prob), online_(online_prob),
          sampling_eval_(sampling_eval), r_order_(seed, 1), r_peer_(seed, 2), r_drop_(seed, 3),
          r_online_(seed, 4), r_delay_(seed, 5), r_eval_(seed, 6), r_token_(seed, 7) {
        if (n_nodes <= 0 || delta <= 0) throw std::invalid_argument("n_nodes and delta must be positive");
        sync_.assign(n_, 1); offset_.assign(n_, 0); round_len_.assign(n_, delta);
        order_.resize(n_);
        for (int i = 0; i < n_; ++i) order_[i] = i;
        accounts_.resize(n_);
    }

    void set_nodes(const std::vector<int>& sync, const std::vector<int>& delta_i, const std::vector<int>& round_len) {
        if ((int)sync.size() != n_ || (int)delta_i.size() != n_ || (int)round_len.size() != n_)
            throw std::invalid_argument("one entry per node expected");
        sync_ = sync; offset_ = delta_i; round_len_ = round_len;
        for (int i = 0; i < n_; ++i)
            if (!sync_[i] && offset_[i] < 1) offset_[i] = 1;     // guard
        queue_valid_ = false;
    }
    void set_topology(const std::vector<int64_t>& indptr, const std::vector<int32_t>& indices) {
        if ((int)indptr.size() != n_ + 1) throw std::invalid_argument("indptr must have n+1 entries");
        indptr_ = indptr; indices_ = indices; clique_ = false;
    }
    void set_delay(int kind, double a, double b) { delay_kind_ = kind; delay_a_ = a; delay_b_ = b; }
    void set_message_sizes(int64_t model_msg, int64_t pull_msg) { size_model_ = model_msg; size_pull_ = pull_msg; }
    void set_token_account(int kind, int64_t C, int64_t A, int64_t k, int64_t utility) {
        for (auto& a : accounts_) { a.kind = kind; a.C = C; a.A = A; a.k = k; a.n = 0; }
        tokenized_ = kind != TK_NONE; utility_ = utility;
    }
    void set_broadcast(bool all_peers) { broadcast_ = all_peers; emit_timeouts_ = all_peers; }   // All2All
    // Restrict node i's peer choice to `peers` (empty = back to its neighbourhood): PENS step 2 gossips only with
    // the peers its selection phase preferred (gossipy/node.py:729-741).  Part of the CONFIGURATION: the caller
    // re-applies it after set_state (the lists are derived from node state that is checkpointed with the nodes).
    void set_peer_list(int i, const std::vector<int32_t>& peers) {
        if (i < 0 || i >= n_) throw std::invalid_argument("node index out of range");
        for (int32_t p : peers)
            if (p < 0 || p >= n_) throw std::invalid_argument("peer index out of range");
        if (peer_list_.empty()) peer_list_.resize(n_);
        peer_list_[i] = peers;
    }

    // Simulate `rounds` rounds starting at the internal clock; returns the events as an int32 array
    // [n_events, 6] = (kind, tick, a, b, slot, aux).
    py::array_t<int32_t> run(int rounds) { return run_ticks((int64_t)rounds * delta_); }
    // Same, for `ticks` timesteps (a round may be simulated in pieces when the caller has to change the
    // configuration at a tick inside it, e.g. the PENS step switch; shuffles and evaluations stay tied to the
    // absolute clock, so run_ticks(a) + run_ticks(b) produce the events of run_ticks(a + b)).
    py::array_t<int32_t> run_ticks(int64_t ticks) {
        if (ticks < 0) throw std::invalid_argument("ticks must be non-negative");
        events_.clear();
        const int64_t first = clock_, last = clock_ + ticks;
        for (int64_t t = first; t < last; ++t) {
            if (t % delta_ == 0) shuffle();
            if (dense_) {
                for (int idx = 0; idx < n_; ++idx) tick_node(order_[idx], t);
            } else {
                fire_due(t);
            }
            online_base_ = r_online_.ctr;                    // this tick's availability draws: one per node, in node order
            r_online_.ctr += (uint64_t)n_;
            deliver_messages(t);
            deliver_replies(t);
            if ((t + 1) % delta_ == 0) evaluate(t);
            clock_ = t + 1;
        }
        const size_t ne = events_.size() / 6;
        py::array_t<int32_t> out({ne, (size_t)6});
        std::copy(events_.begin(), events_.end(), out.mutable_data());
        return out;
    }

    int64_t clock() const { return clock_; }
    void set_clock(int64_t c) { clock_ = c; queue_valid_ = false; }
    // The tick loop of the reference (and of this class until round 2) looks at every node in every tick: N x delta timeout
    // tests and N x delta availability draws per round.  The default loop visits only the nodes that time out at a tick
    // (a queue keyed by their next timeout, processed in the round's shuffled order) and evaluates the availability draw of
    // a node only when a message reaches it (the stream is counter based: draw (tick, node) is addressable).  Same events,
    // same stream counters; `dense = true` keeps the old loop (tests compare the two).
    void set_dense_loop(bool dense) { dense_ = dense; queue_valid_ = false; }
    int64_t sent() const { return sent_; }
    int64_t failed() const { return failed_; }
    int64_t total_size() const { return total_size_; }
    int64_t pending() const {
        int64_t c = 0;
        for (auto& kv : msg_q_) c += (int64_t)kv.second.size();
        for (auto& kv : rep_q_) c += (int64_t)kv.second.size();
        return c;
    }
    // Dynamic state (random-stream counters, round order, token balances, in-flight messages,
    // counters): together with the constructor/set_* configuration it determines every future event,
    // so a checkpointed simulation resumes on exactly the schedule it would have followed.
    py::dict get_state() const {
        py::dict d;
        std::vector<uint64_t> ctr;
        for (const Stream* r : {&r_order_, &r_peer_, &r_drop_, &r_online_, &r_delay_, &r_eval_, &r_token_}) ctr.push_back(r->ctr);
        d["streams"] = ctr;
        d["order"] = order_;
        d["balances"] = token_balances();
        auto dump = [](const std::map<int64_t, std::deque<Msg>>& q) {
            std::vector<std::vector<int64_t>> out;
            for (const auto& kv : q)
                for (const Msg& m : kv.second) out.push_back({kv.first, m.id, m.sender, m.receiver, m.type, m.size});
            return out;
        };
        d["msg_q"] = dump(msg_q_);
        d["rep_q"] = dump(rep_q_);
        d["clock"] = clock_; d["sent"] = sent_; d["failed"] = failed_; d["total_size"] = total_size_;
        d["next_id"] = next_id_;
        d["n_nodes"] = n_;
        return d;
    }
    void set_state(const py::dict& d) {
        if (d["n_nodes"].cast<int>() != n_) throw std::invalid_argument("state belongs to a different node count");
        const auto ctr = d["streams"].cast<std::vector<uint64_t>>();
        Stream* rs[7] = {&r_order_, &r_peer_, &r_drop_, &r_online_, &r_delay_, &r_eval_, &r_token_};
        if (ctr.size() != 7) throw std::invalid_argument("7 stream counters expected");
        for (int i = 0; i < 7; ++i) rs[i]->ctr = ctr[i];
        order_ = d["order"].cast<std::vector<int>>();
        if ((int)order_.size() != n_) throw std::invalid_argument("order must have one entry per node");
        const auto bal = d["balances"].cast<std::vector<int64_t>>();
        for (int i = 0; i < n_ && i < (int)bal.size(); ++i) accounts_[i].n = bal[i];
        auto load = [](std::map<int64_t, std::deque<Msg>>& q, const std::vector<std::vector<int64_t>>& rows) {
            q.clear();
            for (const auto& r : rows) {
                if (r.size() != 6) throw std::invalid_argument("queue rows are (due, id, sender, receiver, type, size)");
                Msg m; m.id = (int32_t)r[1]; m.sender = (int32_t)r[2]; m.receiver = (int32_t)r[3]; m.type = (int32_t)r[4];
                m.size = r[5];
                q[r[0]].push_back(m);
            }
        };
        load(msg_q_, d["msg_q"].cast<std::vector<std::vector<int64_t>>>());
        load(rep_q_, d["rep_q"].cast<std::vector<std::vector<int64_t>>>());
        queue_valid_ = false;
        clock_ = d["clock"].cast<int64_t>(); sent_ = d["sent"].cast<int64_t>(); failed_ = d["failed"].cast<int64_t>();
        total_size_ = d["total_size"].cast<int64_t>(); next_id_ = d["next_id"].cast<int32_t>();
    }

    std::vector<int64_t> token_balances() const {
        std::vector<int64_t> v(n_);
        for (int i = 0; i < n_; ++i) v[i] = accounts_[i].n;
        return v;
    }

private:
    void emit(int32_t kind, int64_t t, int32_t a, int32_t b, int32_t slot, int32_t aux) {
        events_.push_back(kind); events_.push_back((int32_t)t); events_.push_back(a); events_.push_back(b);
        events_.push_back(slot); events_.push_back(aux);
    }
    void shuffle() {   // Fisher-Yates
        for (int i = n_ - 1; i > 0; --i) std::swap(order_[i], order_[(int)r_order_.below((uint64_t)i + 1)]);
        pos_valid_ = false;
    }
    bool is_online(int i) const { return online_ >= 1.0 || r_online_.uniform_at(online_base_ + (uint64_t)i) <= online_; }
    int64_t period_of(int i) const { return sync_[i] ? (int64_t)round_len_[i] : (int64_t)offset_[i]; }
    // first tick >= t at which node i times out (-1: never)
    int64_t first_timeout(int i, int64_t t) const {
        const int64_t p = period_of(i);
        if (p <= 0) return -1;
        if (sync_[i]) {
            if (offset_[i] < 0 || offset_[i] >= p) return -1;
            return t + (((int64_t)offset_[i] - t % p) + p) % p;
        }
        return (t + p - 1) / p * p;
    }
    void rebuild_queue(int64_t t) {
        due_.clear();
        for (int i = 0; i < n_; ++i) {
            const int64_t f = first_timeout(i, t);
            if (f >= 0) due_[f].push_back(i);
        }
        queue_valid_ = true;
    }
    void fire_due(int64_t t) {
        if (!queue_valid_) rebuild_queue(t);
        auto it = due_.find(t);
        if (it == due_.end()) return;
        std::vector<int> nodes = std::move(it->second);
        due_.erase(it);
        if (nodes.size() > 1) {                              // the reference walks the round's shuffled order
            if (!pos_valid_) {
                pos_.resize(n_);
                for (int k = 0; k < n_; ++k) pos_[order_[k]] = k;
                pos_valid_ = true;
            }
            std::sort(nodes.begin(), nodes.end(), [this](int a, int b) { return pos_[a] < pos_[b]; });
        }
        for (int i : nodes) {
            tick_node(i, t);
            due_[t + period_of(i)].push_back(i);
        }
    }
    bool timed_out(int i, int64_t t) const {
        return sync_[i] ? (t % round_len_[i]) == offset_[i] : (t % offset_[i]) == 0;
    }
    int degree(int i) const { return clique_ ? n_ - 1 : (int)(indptr_[i + 1] - indptr_[i]); }
    int peer_at(int i, int k) const {
        if (clique_) return k < i ? k : k + 1;
        return indices_[indptr_[i] + k];
    }
    int64_t delay_of(const Msg& m) {
        switch (delay_kind_) {
            case DELAY_UNIFORM: {
                const int64_t lo = (int64_t)delay_a_, hi = (int64_t)delay_b_;
                return lo + (int64_t)r_delay_.below((uint64_t)(hi - lo + 1));
            }
            case DELAY_LINEAR: return (int64_t)(delay_a_ * (double)m.size) + (int64_t)delay_b_;
            default: return (int64_t)delay_a_;
        }
    }
    void lost(const Msg& m, int64_t t) {
        ++failed_;
        emit(EV_DROP, t, m.sender, m.receiver, m.id, m.type);
    }
    void send_to(int i, int peer, int64_t t) {
        Msg m;
        m.id = next_id_++; m.sender = i; m.receiver = peer;
        m.type = protocol_ == PUSH ? MT_PUSH : (protocol_ == PULL ? MT_PULL : MT_PUSH_PULL);
        m.size = m.type == MT_PULL ? size_pull_ : size_model_;
        emit(EV_SEND, t, i, peer, m.id, m.type);
        ++sent_; total_size_ += m.size;                      // counted at send time (simul.py:401)
        if (r_drop_.uniform() >= drop_) msg_q_[t + delay_of(m)].push_back(m);
        else lost(m, t);
    }
    bool fire(int i, int64_t t) {
        const int deg = degree(i);
        if (deg <= 0) return false;                          // FIX(B6): skip, do not abort the node loop
        if (broadcast_) {
            for (int k = 0; k < deg; ++k) send_to(i, peer_at(i, k), t);
            return true;
        }
        if (!peer_list_.empty() && !peer_list_[i].empty()) {
            const std::vector<int32_t>& pl = peer_list_[i];
            send_to(i, pl[(size_t)r_peer_.below((uint64_t)pl.size())], t);
            return true;
        }
        send_to(i, peer_at(i, (int)r_peer_.below((uint64_t)deg)), t);
        return true;
    }
    void tick_node(int i, int64_t t) {
        if (!timed_out(i, t)) return;
        if (emit_timeouts_) emit(EV_TIMEOUT, t, i, -1, -1, 0);
        if (tokenized_) {
            if (r_token_.uniform() < accounts_[i].proactive()) fire(i, t);
            else { accounts_[i].add(1); emit(EV_TOKEN, t, i, -1, -1, (int32_t)accounts_[i].n); }
        } else {
            fire(i, t);
        }
    }
    void deliver_messages(int64_t t) {
        auto it = msg_q_.find(t);
        if (it == msg_q_.end()) return;
        std::deque<Msg>& q = it->second;                     // may grow while we walk it (delay-0 reactive sends)
        for (size_t k = 0; k < q.size(); ++k) {
            const Msg m = q[k];
            if (!is_online(m.receiver)) { lost(m, t); continue; }
            emit(EV_DELIVER, t, m.sender, m.receiver, m.id, m.type);
            const bool wants_reply = m.type == MT_PULL || m.type == MT_PUSH_PULL;
            if (wants_reply) {
                Msg r;
                r.id = next_id_++; r.sender = m.receiver; r.receiver = m.sender; r.type = MT_REPLY;
                r.size = size_model_;
                emit(EV_REPLY_SEND, t, m.sender, m.receiver, m.id, r.id);
                if (r_drop_.uniform() > drop_) rep_q_[t + delay_of(r)].push_back(r);   // `>` for replies (simul.py:414)
                else lost(r, t);
            } else if (tokenized_) {                         // the RECEIVER reacts (FIX B4/B5)
                TokenAccount& acc = accounts_[m.receiver];
                const int64_t reaction = acc.reactive(utility_, r_token_);
                if (reaction > 0) {
                    acc.sub(reaction);
                    for (int64_t c = 0; c < reaction; ++c)
                        if (!fire(m.receiver, t)) break;
                }
            }
        }
        msg_q_.erase(t);
    }
    void deliver_replies(int64_t t) {
        auto it = rep_q_.find(t);
        if (it == rep_q_.end()) return;
        for (const Msg& r : it->second) {
            if (is_online(r.receiver)) {
                ++sent_; total_size_ += r.size;              // replies are counted at delivery (simul.py:425)
                emit(EV_REPLY_DELIVER, t, r.receiver, r.sender, r.id, r.type);
            } else {
                lost(r, t);
            }
        }
        rep_q_.erase(it);
    }
    void evaluate(int64_t t) {
        if (sampling_eval_ > 0) {
            const int k = std::max((int)(n_ * sampling_eval_), 1);
            for (int c = 0; c < k; ++c) emit(EV_EVAL, t, (int32_t)r_eval_.below((uint64_t)n_), -1, -1, 0);   // with replacement
        } else {
            for (int i = 0; i < n_; ++i) emit(EV_EVAL, t, i, -1, -1, 0);
        }
    }

    int n_, delta_, protocol_;
    double drop_, online_, sampling_eval_;
    Stream r_order_, r_peer_, r_drop_, r_online_, r_delay_, r_eval_, r_token_;
    std::vector<int> sync_, offset_, round_len_, order_;
    std::vector<int64_t> indptr_; std::vector<int32_t> indices_; bool clique_ = true;
    std::vector<std::vector<int32_t>> peer_list_;          // per-node restriction of the peer choice (empty = none)
    int delay_kind_ = DELAY_CONST; double delay_a_ = 0, delay_b_ = 0;
    int64_t size_model_ = 1, size_pull_ = 1;
    std::vector<TokenAccount> accounts_; bool tokenized_ = false; int64_t utility_ = 1;
    bool broadcast_ = false, emit_timeouts_ = false;
    uint64_t online_base_ = 0;                             // counter of this tick's first availability draw
    bool dense_ = false, queue_valid_ = false, pos_valid_ = false;
    std::map<int64_t, std::vector<int>> due_;              // next timeout -> nodes
    std::vector<int> pos_;                                 // position of every node in the round's shuffled order
    std::map<int64_t, std::deque<Msg>> msg_q_, rep_q_;
    std::vector<int32_t> events_;
    int64_t clock_ = 0, sent_ = 0, failed_ = 0, total_size_ = 0;
    int32_t next_id_ = 0;
};

void bind_scheduler(py::module_& m) {
    py::class_<GossipScheduler>(m, "GossipScheduler",
                                "Native control plane of the gossip round loop (see csrc/sched/scheduler.cpp)")
        .def(py::init<int, int, int, double, double, double, uint64_t>(), py::arg("n_nodes"), py::arg("delta"),
             py::arg("protocol"), py::arg("drop_prob") = 0.0, py::arg("online_prob") = 1.0,
             py::arg("sampling_eval") = 0.0, py::arg("seed") = 0)
        .def("set_nodes", &GossipScheduler::set_nodes)
        .def("set_topology", &GossipScheduler::set_topology)
        .def("set_delay", &GossipScheduler::set_delay)
        .def("set_message_sizes", &GossipScheduler::set_message_sizes)
        .def("set_token_account", &GossipScheduler::set_token_account)
        .def("set_broadcast", &GossipScheduler::set_broadcast)
        .def("set_peer_list", &GossipScheduler::set_peer_list)
        .def("set_dense_loop", &GossipScheduler::set_dense_loop)
        .def("run", &GossipScheduler::run, py::arg("rounds") = 1)
        .def("run_ticks", &GossipScheduler::run_ticks, py::arg("ticks"))
        .def_property("clock", &GossipScheduler::clock, &GossipScheduler::set_clock)
        .def_property_readonly("sent", &GossipScheduler::sent)
        .def_property_readonly("failed", &GossipScheduler::failed)
        .def_property_readonly("total_size", &GossipScheduler::total_size)
        .def_property_readonly("pending", &GossipScheduler::pending)
        .def("token_balances", &GossipScheduler::token_balances)
        .def("get_state", &GossipScheduler::get_state)
        .def("set_state", &GossipScheduler::set_state);
    m.attr("EV_SEND") = (int)EV_SEND; m.attr("EV_DROP") = (int)EV_DROP; m.attr("EV_DELIVER") = (int)EV_DELIVER;
    m.attr("EV_REPLY_SEND") = (int)EV_REPLY_SEND; m.attr("EV_REPLY_DELIVER") = (int)EV_REPLY_DELIVER;
    m.attr("EV_EVAL") = (int)EV_EVAL; m.attr("EV_TOKEN") = (int)EV_TOKEN; m.attr("EV_TIMEOUT") = (int)EV_TIMEOUT;
}

}  // namespace gb
