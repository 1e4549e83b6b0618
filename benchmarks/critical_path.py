#!/usr/bin/env python
"""Dependency-chain bound of the headline benchmark (runs on CPU, no GPU needed).

A gossip round is not a bag of independent updates: a node's update waits for the sender's snapshot,
the reply carries the *updated* model, and everything a node does is ordered on its stream.  This
script replays the native scheduler's event list of the bench configuration (8 nodes, PUSH_PULL,
delta 100) through a small list-scheduling model -- one in-order stream per node, unlimited SMs, a
host that needs ``--host-us`` per enqueued kernel -- and prints the makespan per round for the
measured kernel durations.  That number is the "speed of light" of the round *given* the kernels:
the gap between it and ``bench.py``'s ms/round is what a better executor could still recover, the
bound itself only moves with a faster local-update kernel.

    python benchmarks/critical_path.py [--rounds 200] [--update-us 1090] [--host-us 0 10 25 50]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def events_of(n_nodes: int, delta: int, rounds: int, seed: int = 1234):
    from gossipy_b200.ops.native import _try_import
    C = _try_import()
    sch = C.GossipScheduler(n_nodes, delta, 3, 0.0, 1.0, 0.0, seed)       # 3 = PUSH_PULL
    import random
    rnd = random.Random(seed)
    sch.set_nodes([1] * n_nodes, [rnd.randrange(delta) for _ in range(n_nodes)], [delta] * n_nodes)   # sync nodes: fixed offset in the round
    indptr = [i * (n_nodes - 1) for i in range(n_nodes + 1)]
    indices = [j for i in range(n_nodes) for j in range(n_nodes) if j != i]
    sch.set_topology(indptr, indices)
    sch.set_delay(0, 0.0, 0.0)
    sch.set_message_sizes(79511, 1)
    return C, [sch.run(1).tolist() for _ in range(rounds)]


def makespan(C, rounds_events, n_nodes, t_upd, t_snap, t_eval, host_us, host_fixed_us=0.0):
    ready = [0.0] * n_nodes          # per-node stream
    msg = {}                         # slot -> time the snapshot is complete
    host = 0.0
    ends = []
    n_upd = 0
    for ev in rounds_events:
        host += host_fixed_us        # scheduler call, event decoding, metric bookkeeping
        for kind, t, a, b, slot, aux in ev:
            if kind == C.EV_SEND:
                host += host_us
                s = max(ready[a], host)
                ready[a] = s + t_snap
                msg[slot] = ready[a]
            elif kind == C.EV_DELIVER:
                host += host_us
                s = max(ready[b], msg.pop(slot), host)
                ready[b] = s + t_upd
                n_upd += 1
                last = b
            elif kind == C.EV_REPLY_SEND:
                host += host_us
                s = max(ready[last], host)
                ready[last] = s + t_snap
                msg[aux] = ready[last]
            elif kind == C.EV_REPLY_DELIVER:
                host += host_us
                s = max(ready[a], msg.pop(slot), host)
                ready[a] = s + t_upd
                n_upd += 1
            elif kind == C.EV_EVAL:
                host += host_us
                s = max(ready[a], host)
                ready[a] = s + t_eval
        ends.append(max(ready))
    return ends, n_upd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=8)
    ap.add_argument("--delta", type=int, default=100)
    ap.add_argument("--rounds", type=int, default=200)
    ap.add_argument("--update-us", type=float, nargs="*", default=[1087.0, 1290.0],
                    help="merge + local epoch incl. the loader launch (1 050 + 37 us alone; 1 290 us with 8 running concurrently)")
    ap.add_argument("--snap-us", type=float, default=8.3)
    ap.add_argument("--eval-us", type=float, default=27.7)
    ap.add_argument("--host-us", type=float, nargs="*", default=[0.0, 25.0, 100.0])
    ap.add_argument("--seeds", type=int, nargs="*", default=[1, 2, 3, 4, 5, 6, 7, 8])
    args = ap.parse_args()
    skip = args.rounds // 10
    for tu in args.update_us:
        for h in args.host_us:
            per_seed = []
            for seed in args.seeds:
                C, evs = events_of(args.nodes, args.delta, args.rounds, seed)
                ends, n_upd = makespan(C, evs, args.nodes, tu, args.snap_us, args.eval_us, h)
                per_seed.append((ends[-1] - ends[skip - 1]) / (len(ends) - skip))
            per_seed.sort()
            med = per_seed[len(per_seed) // 2]
            print(json.dumps({"update_us": tu, "host_us_per_launch": h, "updates_per_round": n_upd / len(evs),
                              "bound_ms_per_round": {"min": round(per_seed[0] / 1e3, 3), "median": round(med / 1e3, 3),
                                                     "max": round(per_seed[-1] / 1e3, 3)},
                              "bound_rounds_per_s_median": round(1e6 / med, 1),
                              "chain_updates_per_round_median": round(med / tu, 2)}))


if __name__ == "__main__":
    main()
